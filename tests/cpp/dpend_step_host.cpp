// Host-side check of csrc/dpend_step.cuh (the header compiles for the host with the same fma() sequence as the kernel):
// the carried-pair RK4 step of the double pendulum against a long-double RK4 of the reference's equations
// (physics/pendulum/double_pendulum.hpp:126-179, core/solver.hpp:95-125), step by step FROM THE SAME INPUTS along whole
// trajectories, so that the rounding carried in the pairs between two refreshes is part of what is measured.  Beside it the
// plain-double RK4 in the reference's own association against the same truth: the step must stay in that error class.
// Ensembles: the BASELINE one (m = L = 1, g = 9.81, theta ~ U(-3, 3)), jittered parameters, g = 0, m2 = 0, fast spinners and
// coarse steps that drive every fall-back tier; sbdp::step_optimistic (the one-branch form) must reproduce sbdp::step bit for bit.  Pure CPU; run by tests/test_oracle.py.  Prints one line per ensemble:
//   name  max_rel_err_fast  max_rel_err_plain  tier counts
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>
#include "../../sopot_b200/csrc/dpend_step.cuh"

struct StepK { double dt, hdt, dt6, hdt2, rot3, rot5, cot4, dt2_6; };     // = StepConsts of csrc/ensemble.cuh
static StepK make_k(double dt) { return StepK{dt, 0.5 * dt, dt / 6.0, (0.5 * dt) * (0.5 * dt), -1.0 / 6.0, 1.0 / 120.0, 1.0 / 24.0, dt * dt / 6.0}; }

struct Par { double m1, m2, L1, L2, g; };

template <class T>
static void rhs(const Par& p, const T (&y)[4], T (&dy)[4]) {
    const T delta = y[0] - y[1];
    const T sd = std::sin(delta), cd = std::cos(delta), s1 = std::sin(y[0]), s2 = std::sin(y[1]);
    const T a11 = (T(p.m1) + T(p.m2)) * T(p.L1), a12 = T(p.m2) * T(p.L2) * cd, a21 = T(p.L1) * cd, a22 = T(p.L2);
    const T b1 = T(p.m2) * T(p.L2) * y[3] * y[3] * sd - (T(p.m1) + T(p.m2)) * T(p.g) * s1;
    const T b2 = T(-p.L1) * y[2] * y[2] * sd - T(p.g) * s2;
    const T det = a11 * a22 - a12 * a21;
    dy[0] = y[2]; dy[1] = y[3];
    dy[2] = (b1 * a22 - b2 * a12) / det;
    dy[3] = (a11 * b2 - a21 * b1) / det;
}
template <class T>
static void rk4(const Par& p, T (&y)[4], T dt) {
    T k1[4], k2[4], k3[4], k4[4], t[4];
    rhs(p, y, k1); for (int i = 0; i < 4; ++i) { k1[i] *= dt; t[i] = y[i] + T(0.5) * k1[i]; }
    rhs(p, t, k2); for (int i = 0; i < 4; ++i) { k2[i] *= dt; t[i] = y[i] + T(0.5) * k2[i]; }
    rhs(p, t, k3); for (int i = 0; i < 4; ++i) { k3[i] *= dt; t[i] = y[i] + k3[i]; }
    rhs(p, t, k4); for (int i = 0; i < 4; ++i) { k4[i] *= dt; y[i] = y[i] + (k1[i] + T(2.0) * k2[i] + T(2.0) * k3[i] + k4[i]) / T(6.0); }
}

struct Result { double fast = 0, plain = 0; long half = 0, full = 0, steps = 0; };

static void run(const char* name, int n_inst, int n_steps, double dt, unsigned seed, int kind, double tol) {
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    const StepK k = make_k(dt);
    Result r;
    for (int i = 0; i < n_inst; ++i) {
        Par p{1.0, 1.0, 1.0, 1.0, 9.81};
        double y[4] = {3.0 * U(rng), 3.0 * U(rng), 0.0, 0.0};
        if (kind == 1) { p = Par{1.0 + 0.5 * U(rng), 1.0 + 0.5 * U(rng), 1.0 + 0.5 * U(rng), 1.0 + 0.5 * U(rng), 9.81 * (1.0 + 0.1 * U(rng))}; y[2] = 2.0 * U(rng); y[3] = 2.0 * U(rng); }
        if (kind == 2) { p.g = 0.0; y[2] = 3.0 * U(rng); y[3] = 3.0 * U(rng); }
        if (kind == 3) { p.m2 = 0.0; y[2] = U(rng); y[3] = U(rng); }
        if (kind == 4) { y[2] = 40.0 * U(rng); y[3] = 40.0 * U(rng); }                  // beyond the 2^-6 tiers, partly beyond 2^-5
        if (kind == 5) { y[0] += 1000.0 * std::round(3.0 * U(rng)); y[2] = 5.0 * U(rng); }   // angles far from zero
        sbdp::Scales q;
        q.lam = p.L1 / p.L2;
        q.kol = ((p.m2 * p.L2) / ((p.m1 + p.m2) * p.L1)) / q.lam;
        q.gL1 = p.g / p.L1;
        sbdp::Pairs b;
        sbdp::pairs_of<false>(q, y[0], y[0] - y[1], b);
        for (int s = 0; s < n_steps; ++s) {
            long double yt[4] = {y[0], y[1], y[2], y[3]};
            rk4<long double>(p, yt, (long double)dt);
            double yp[4] = {y[0], y[1], y[2], y[3]};
            rk4<double>(p, yp, dt);
            const double x1 = dt * y[2], xd = x1 - dt * y[3];
            r.half += sbdp::below_2m6(x1) && sbdp::below_2m6(xd);
            // the one-branch form of the step must give the same bits from the same inputs
            double yo[4] = {y[0], y[1], y[2], y[3]};
            sbdp::Pairs bo = b;
            sbdp::step_optimistic(q, bo, yo, k, ((s + 1) & 15) == 0);
            sbdp::step(q, b, y, k, ((s + 1) & 15) == 0);
            if (std::memcmp(yo, y, sizeof y) != 0 || std::memcmp(&bo, &b, sizeof b) != 0) {
                std::printf("%s: step_optimistic differs from step at instance %d step %d\n", name, i, s);
                std::exit(1);
            }
            r.steps++;
            for (int d = 0; d < 4; ++d) {
                // relative to the magnitude of the component, angles and rates floored at 1 (an angle or a rate passing through
                // zero has no relative error to speak of)
                const long double scale = std::fmax(1.0, (double)std::fabs(yt[d]));
                r.fast = std::fmax(r.fast, (double)(std::fabs((long double)y[d] - yt[d]) / scale));
                r.plain = std::fmax(r.plain, (double)(std::fabs((long double)yp[d] - yt[d]) / scale));
            }
        }
    }
    std::printf("%-10s dt %.0e  fast %.3e  plain %.3e  first-tier steps %.4f  %s\n", name, dt, r.fast, r.plain, (double)r.half / (double)r.steps,
                r.fast <= tol ? "ok" : "FAIL");
    if (!(r.fast <= tol)) std::exit(1);
}

// the two tiered rotations against the addition theorem in long double, absolute error in ulp(1) of a unit pair
static void check_rotations(long n) {
    std::mt19937_64 rng(11);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    double worst_half = 0, worst_full = 0;
    for (long i = 0; i < n; ++i) {
        const long double x = 3.5L * U(rng), sl = sinl(x), cl = cosl(x);
        // every 8th sample on the bound of its tier
        const double xh = (i & 7) ? sbdp::kHalfTier * U(rng) : ((i & 8) ? sbdp::kHalfTier : -sbdp::kHalfTier);
        const double hf = (i & 7) ? sbdp::kFullTier * U(rng) : ((i & 8) ? sbdp::kFullTier : -sbdp::kFullTier);
        if (!sbdp::below_2m6(xh) || !sbdp::below_2m6(hf)) { std::puts("below_2m6 rejects a value inside the tier"); std::exit(1); }
        double rs, rc;
        sbdp::rotate_half((double)sl, (double)cl, xh, rs, rc);
        long double sh = sinl(0.5L * xh), ch = cosl(0.5L * xh);
        const long double s0 = (double)sl, c0 = (double)cl;      // the rounded pair is what gets rotated
        worst_half = std::fmax(worst_half, (double)std::fabs((long double)rs - (s0 * ch + c0 * sh)) / 0x1p-52);
        worst_half = std::fmax(worst_half, (double)std::fabs((long double)rc - (c0 * ch - s0 * sh)) / 0x1p-52);
        sbdp::rotate_2m6((double)sl, (double)cl, hf, -1.0 / 6.0, 1.0 / 24.0, rs, rc);
        sh = sinl((long double)hf); ch = cosl((long double)hf);
        worst_full = std::fmax(worst_full, (double)std::fabs((long double)rs - (s0 * ch + c0 * sh)) / 0x1p-52);
        worst_full = std::fmax(worst_full, (double)std::fabs((long double)rc - (c0 * ch - s0 * sh)) / 0x1p-52);
    }
    if (sbdp::below_2m6(0x1.0001p-6) || sbdp::below_2m6(-0.02)) { std::puts("below_2m6 accepts a value outside the tier"); std::exit(1); }
    std::printf("rotate_half max %.3f ulp(1)  rotate_2m6 max %.3f ulp(1)  %s\n", worst_half, worst_full,
                worst_half <= 1.0 && worst_full <= 1.0 ? "ok" : "FAIL");
    if (!(worst_half <= 1.0 && worst_full <= 1.0)) std::exit(1);
}

// integrate: dpend_step_host integrate <in.bin> <out.bin> <n> <steps> <dt> <store_every>
// in.bin = 9 arrays of n doubles (m1, m2, L1, L2, g, th1, th2, w1, w2), out.bin = [snapshot][4][n] -- the kernel's loop
// (load, refresh every 16th step, snapshots) on the host, for the full-horizon comparison with the oracle in tests/test_oracle.py
static int integrate(int argc, char** argv) {
    if (argc < 8) return 2;
    const long n = std::atol(argv[4]), steps = std::atol(argv[5]), every = std::atol(argv[7]);
    const double dt = std::atof(argv[6]);
    std::FILE* f = std::fopen(argv[2], "rb");
    if (!f) return 2;
    std::vector<double> in(9 * n);
    if (std::fread(in.data(), 8, in.size(), f) != in.size()) return 2;
    std::fclose(f);
    const long snaps = every ? steps / every : 0;
    std::vector<double> out((snaps + 1) * 4 * n);
    const StepK k = make_k(dt);
    for (long i = 0; i < n; ++i) {
        const double m1 = in[i], m2 = in[n + i], L1 = in[2 * n + i], L2 = in[3 * n + i], g = in[4 * n + i];
        double y[4] = {in[5 * n + i], in[6 * n + i], in[7 * n + i], in[8 * n + i]};
        sbdp::Scales q;
        q.lam = L1 / L2;
        q.kol = ((m2 * L2) / ((m1 + m2) * L1)) / q.lam;
        q.gL1 = g / L1;
        sbdp::Pairs b;
        sbdp::pairs_of<false>(q, y[0], y[0] - y[1], b);
        long snap = 0;
        for (long s = 0; s < steps; ++s) {
            sbdp::step(q, b, y, k, ((s + 1) & 15) == 0);
            if (every && (s + 1) % every == 0) { for (int d = 0; d < 4; ++d) out[(snap * 4 + d) * n + i] = y[d]; ++snap; }
        }
        for (int d = 0; d < 4; ++d) out[(snaps * 4 + d) * n + i] = y[d];
    }
    f = std::fopen(argv[3], "wb");
    if (!f || std::fwrite(out.data(), 8, out.size(), f) != out.size()) return 2;
    std::fclose(f);
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 1 && std::string(argv[1]) == "integrate") return integrate(argc, argv);
    const int scale = argc > 1 ? std::atoi(argv[1]) : 1;
    check_rotations(400000L * scale);
    // tolerance: 1e-12 relative per step is the mode's promise; the step is expected (and required here) to stay within a few ulp
    run("baseline", 64 * scale, 10000, 1e-3, 2, 0, 4e-15);
    run("jittered", 64 * scale, 4000, 1e-3, 3, 1, 4e-15);
    run("g=0", 16 * scale, 4000, 1e-3, 4, 2, 4e-15);
    run("m2=0", 16 * scale, 4000, 1e-3, 5, 3, 4e-15);
    run("spinners", 32 * scale, 2000, 1e-3, 6, 4, 1e-13);
    run("far-angle", 16 * scale, 2000, 1e-3, 7, 5, 4e-13);          // ulp(1000 rad) = 1.1e-13 enters every sin/cos argument
    run("coarse", 32 * scale, 1000, 1e-2, 8, 0, 2e-14);
    run("fine", 16 * scale, 4000, 1e-4, 9, 0, 4e-15);
    return 0;
}
