// dpend_step.cuh -- the FP_CONTRACT RK4 step of the double pendulum (physics/pendulum/double_pendulum.hpp:126-179 under
// core/solver.hpp:95-125), third form (SB_DPEND_STEP = 3).  Host-compilable like trig.cuh: tests/cpp/dpend_step_host.cpp runs
// the very same fma() sequence on the CPU against a long-double RK4 of the reference's equations.
//
// Same carried, scaled pairs as the second form (ho_dpend.cu):  (s1, c1) = (g/L1)(sin th1, cos th1),  (S, C) = lam (sin, cos)(th1 - th2),
// kol = kap / lam, lam = L1/L2, kap = m2 L2 / ((m1+m2) L1).  What changed, in executed FP64 instructions per instance-step
// (ncu: 174.0 -> 158.2; 45.57 -> 41.91 ms per 1000 steps of 4 Mi instances, profiles/r2_dpend_bench_step3.txt):
//   * derivative 16 -> 15:  -B2 = lam w1^2 sd + (g/L2) sin th2 = S (w1^2 - c1) + s1 C   [(g/L2) sin th2 = s1 C - c1 S]
//   * the increments of the difference angle by one FMA from the increment of th1 (h_d = h_1 - c w2) instead of (w1 - w2) c;
//   * stage 2 rotates by HALF the full-step increment x = dt w, which the Nystrom angle row needs anyway (polynomials in x
//     with the powers of 1/2 folded into the coefficients), so dt/2 w is never formed;
//   * tiered rotation polynomials: stage 2, |x/2| <= 2^-7: sin to h^5 (next term h^7/5040 <= 3.5e-19), cos - 1 = h^2 (-1/2 + b h^2)
//     with an economised b (error <= 3.3e-17 = 0.15 ulp(1)); stage 4, |h| <= 2^-6: sin to h^5 (<= 4.5e-17 = 0.2 ulp(1) at
//     the bound, 0.01 ulp at |h| = 2^-6.6), cos to h^6.  Larger increments take the second form's rotate_small_k (|h| <= 2^-5),
//     then the full sincos.  On the BASELINE ensemble 0.4 % of the warps hold a lane beyond the first tier at any time
//     (|w| <= 15.6 rad/s; measured with the oracle);
//   * one running sum A = a_2 + a_3 per rate serves both the Nystrom angle row (a_1 + A) and the rate row (a_1 + 2A + a_4).
#pragma once
#include "trig.cuh"

namespace sbdp {

// Every polynomial FMA below has ONE coefficient that is not an instruction immediate: an FP64 instruction encodes a single
// constant-bank / uniform operand, so of two full-precision coefficients one ends up in a vector register and the FMA reads
// three of them (3 issue cycles, profiles/r1_fp64_operand_rule.txt).  Coefficients of the highest power are therefore
// truncated to their high word (relative 6e-8 on a term below 2^-29 of the result), -1/2 is exact, and the economised cosine
// carries its whole correction in the h^4 coefficient:
//   cos h - 1 = h^2 (-1/2 + kC4 h^2),  kC4 = 1/24 - beta Z/720,  Z = 2^-14,  beta = 0.8941 (root of 0.14815 b^3 + b - 1: equal
//   error at the interior extremum and at h^2 = Z)  ->  |error| <= (1 - beta) Z^3/720 = 3.3e-17 = 0.15 ulp(1) on |h| <= 2^-7.
constexpr double kC4 = 1.0 / 24.0 - 0.8941074569749823 * 0x1p-14 / 720.0;
constexpr double kT5hi = 0x1.1111100000000p-7;                          // 1/120 truncated
// stage 2 in terms of x = 2h:  sin h = x (1/2 + x^2 (kH3 + x^2 kH5hi)),  cos h = 1 + x^2 (-1/8 + x^2 kG4)
constexpr double kH3 = -1.0 / 48.0, kH5hi = 0x1.1111100000000p-12, kG4 = kC4 / 16.0;
constexpr double kHalfTier = 0x1p-6;        // |x| <= 2^-6  <=>  |h| <= 2^-7
constexpr double kFullTier = 0x1p-6;        // stage 4: |h| <= 2^-6

#ifdef __CUDACC__
__constant__ double c_dp3[2] = {kH3, kG4};
#endif
constexpr double h_dp3[2] = {kH3, kG4};
#ifdef __CUDA_ARCH__
#define SB_DP3(i) sbdp::c_dp3[i]
#else
#define SB_DP3(i) sbdp::h_dp3[i]
#endif

SB_HD bool below_2m6(double h) {
#ifdef __CUDA_ARCH__
    return (__double2hiint(h) & 0x7fffffff) <= 0x3f900000;    // |h| <= 2^-6 (up to the low word), integer compare
#else
    uint64_t u; std::memcpy(&u, &h, 8); return (uint32_t)((u >> 32) & 0x7fffffff) <= 0x3f900000u;
#endif
}

// rotate (s, c) by x/2, |x| <= 2^-6: 10 FP64 instructions
SB_HD void rotate_half(double s, double c, double x, double& s_out, double& c_out) {
    const double x2 = x * x;
    const double p = fma(x2, kH5hi, SB_DP3(0));
    const double sh = x * fma(x2, p, 0.5);
    const double q = fma(x2, SB_DP3(1), -0.125);
    const double ch = fma(x2, q, 1.0);
    s_out = fma(s, ch, c * sh);
    c_out = fma(c, ch, -(s * sh));
}
// rotate (s, c) by h, |h| <= 2^-6: 11 FP64 instructions (T3, U4 from the caller's constant-bank operands)
SB_HD void rotate_2m6(double s, double c, double h, double T3, double U4, double& s_out, double& c_out) {
    const double h2 = h * h;
    const double p = fma(h2, kT5hi, T3);
    const double sh = h * fma(h2, p, 1.0);
    double q = fma(h2, sbtrig::kU6hi, U4); q = fma(h2, q, -0.5);
    const double ch = fma(h2, q, 1.0);
    s_out = fma(s, ch, c * sh);
    c_out = fma(c, ch, -(s * sh));
}

struct Pairs { double s1, c1, S, C; };            // (g/L1)(sin, cos) th1,  lam (sin, cos)(th1 - th2)
struct Scales { double kol, lam, gL1; };          // kap/lam, L1/L2, g/L1

// the row-scaled Cramer solve of :162-175 on the carried pairs: 15 FP64 instructions + 1 MUFU, 4 three-register FMAs
SB_HD void accel(const double kol, const Pairs& p, const double w1, const double w2, double& a1, double& a2) {
    const double t = fma(w1, w1, -p.c1);                          // w1^2 - (g/L1) cos th1
    const double nB2 = fma(p.s1, p.C, p.S * t);                   // lam w1^2 sd + (g/L2) sin th2 = -B2
    const double kc = kol * p.C;                                  // kap cd
    const double B1 = fma(kol * p.S, w2 * w2, -p.s1);             // kap w2^2 sd - (g/L1) sin th1
    const double D = fma(-kc, p.C, 1.0);                          // 1 - kap lam cd^2 in (0, 1]
    const double inv = sbtrig::rcp_fast(D);
    a1 = fma(kc, nB2, B1) * inv;
    a2 = -fma(p.C, B1, nB2) * inv;
}

template <bool COLD>
SB_HD void pairs_of(const Scales& q, double th1, double delta, Pairs& p) {
    double st, ct, sd, cd;
    sbtrig::sincos<COLD>(th1, st, ct);
    sbtrig::sincos<COLD>(delta, sd, cd);
    p.s1 = q.gL1 * st; p.c1 = q.gL1 * ct;
    p.S = q.lam * sd; p.C = q.lam * cd;
}

// pairs at (th1 + h1, delta + hd) from the base pairs by the second form's small rotation, else from the angles
template <class K>
SB_HD void rotate_any(const Scales& q, const K& k, double th1, double th2, const Pairs& b, double h1, double hd, Pairs& o) {
    if (sbtrig::small_angle(h1) && sbtrig::small_angle(hd)) {
        sbtrig::rotate_small_k(b.s1, b.c1, h1, k.rot3, k.rot5, k.cot4, o.s1, o.c1);
        sbtrig::rotate_small_k(b.S, b.C, hd, k.rot3, k.rot5, k.cot4, o.S, o.C);
    } else {
        pairs_of<true>(q, th1 + h1, (th1 - th2) + hd, o);
    }
}

// One RK4 step  y += (dt/6)(f1 + 2 f2 + 2 f3 + f4)  of y = (th1, th2, w1, w2); `b` holds the pairs of (th1, th2) on entry and
// of the new angles on exit.  K: the step constants {dt, hdt, dt6, hdt2, rot3, rot5, cot4, dt2_6} (ensemble.cuh).
template <class K>
SB_HD void step(const Scales& q, Pairs& b, double (&y)[4], const K& k, const bool refresh) {
    const double th1 = y[0], th2 = y[1], w1 = y[2], w2 = y[3];
    const double x1 = k.dt * w1, x2 = k.dt * w2, xd = x1 - x2;                    // full-step increments by the base rates
    double a1, a2;
    accel(q.kol, b, w1, w2, a1, a2);                                              // f1
    const double a1f = a1, a2f = a2;
    double u1 = fma(k.hdt, a1, w1), u2 = fma(k.hdt, a2, w2);                      // stage-2 rates
    Pairs p;
    if (below_2m6(x1) && below_2m6(xd)) {
        rotate_half(b.s1, b.c1, x1, p.s1, p.c1);
        rotate_half(b.S, b.C, xd, p.S, p.C);
    } else {
        rotate_any(q, k, th1, th2, b, 0.5 * x1, 0.5 * xd, p);
    }
    accel(q.kol, p, u1, u2, a1, a2);                                              // f2
    double A1 = a1, A2 = a2;
    const double d1 = k.hdt2 * a1f, dd = fma(-k.hdt2, a2f, d1);                   // stage 3 - stage 2 angles: (dt/2)^2 a_1
    if (sbtrig::tiny_angle(d1) && sbtrig::tiny_angle(dd)) {
        sbtrig::rotate_tiny_k(p.s1, p.c1, d1, k.rot3, p.s1, p.c1);
        sbtrig::rotate_tiny_k(p.S, p.C, dd, k.rot3, p.S, p.C);
    } else {
        rotate_any(q, k, th1, th2, b, k.hdt * u1, k.hdt * (u1 - u2), p);
    }
    u1 = fma(k.hdt, a1, w1); u2 = fma(k.hdt, a2, w2);                             // stage-3 rates
    accel(q.kol, p, u1, u2, a1, a2);                                              // f3
    A1 += a1; A2 += a2;
    const double h41 = k.dt * u1, h4d = fma(-k.dt, u2, h41);                      // stage-4 angles - base angles
    if (below_2m6(h41) && below_2m6(h4d)) {
        rotate_2m6(b.s1, b.c1, h41, k.rot3, k.cot4, p.s1, p.c1);
        rotate_2m6(b.S, b.C, h4d, k.rot3, k.cot4, p.S, p.C);
    } else {
        rotate_any(q, k, th1, th2, b, h41, h4d, p);
    }
    u1 = fma(k.dt, a1, w1); u2 = fma(k.dt, a2, w2);                               // stage-4 rates
    accel(q.kol, p, u1, u2, a1, a2);                                              // f4
    // Nystrom form of the angle rows: g = dt w + dt^2/6 (a_1 + a_2 + a_3); rate rows: w += dt/6 (a_1 + 2 (a_2 + a_3) + a_4)
    const double g1 = fma(k.dt2_6, a1f + A1, x1), g2 = fma(k.dt2_6, a2f + A2, x2);
    const double n1 = th1 + g1, n2 = th2 + g2;
    y[0] = n1; y[1] = n2;
    y[2] = fma(k.dt6, fma(2.0, A1, a1f + a1), w1);
    y[3] = fma(k.dt6, fma(2.0, A2, a2f + a2), w2);
    // new base angle - stage-4 angle = dt^2/6 (a_1 - 2 a_2 + a_3) = O(dt^3): micro rotation of the stage-4 pairs
    const double gd = g1 - g2;
    const double e1 = g1 - h41, ed = gd - h4d;
    if (!refresh && sbtrig::micro_angle(e1) && sbtrig::micro_angle(ed)) {
        sbtrig::rotate_micro(p.s1, p.c1, e1, b.s1, b.c1);
        sbtrig::rotate_micro(p.S, p.C, ed, b.S, b.C);
    } else if (!refresh && sbtrig::tiny_angle(e1) && sbtrig::tiny_angle(ed)) {
        sbtrig::rotate_tiny_k(p.s1, p.c1, e1, k.rot3, b.s1, b.c1);
        sbtrig::rotate_tiny_k(p.S, p.C, ed, k.rot3, b.S, b.C);
    } else if (!refresh && sbtrig::small_angle(g1) && sbtrig::small_angle(gd)) {
        const Pairs o = b;
        sbtrig::rotate_small_k(o.s1, o.c1, g1, k.rot3, k.rot5, k.cot4, b.s1, b.c1);
        sbtrig::rotate_small_k(o.S, o.C, gd, k.rot3, k.rot5, k.cot4, b.S, b.C);
    } else {
        pairs_of<true>(q, n1, n1 - n2, b);                                        // delta rounded as in the reference (:145)
    }
}

// |x| as an ordered integer (high word without the sign): lets several range tests share one compare
SB_HD int hi_abs(double x) {
#ifdef __CUDA_ARCH__
    return __double2hiint(x) & 0x7fffffff;
#else
    uint64_t u; std::memcpy(&u, &x, 8); return (int)((u >> 32) & 0x7fffffff);
#endif
}
SB_HD int imax(int a, int b) { return a > b ? a : b; }

// out-of-line copy of step() for the second pass (keeps the registers of the rare route out of the time loop's allocation)
#ifdef __CUDACC__
#define SB_NOINLINE __host__ __device__ __noinline__
#else
#define SB_NOINLINE __attribute__((noinline))
#endif
template <class K>
SB_NOINLINE void step_out_of_line(const Scales& q, Pairs& b, double (&y)[4], const K& k, const bool refresh) { step(q, b, y, k, refresh); }

// The same step with ONE branch: the first-tier route of step() is computed unconditionally, the eight range tests are folded
// into three integer maxima, and a step in which any increment leaves its tier is recomputed by step() from the saved
// inputs -- bit-identical to step() in every case (on the BASELINE ensemble 0.4 % of the warp-steps take the second pass).
// Removes the four divergence points (BSSY / BSYNC / BRA, 27 control instructions) from the hot path of a step -- and measured
// 3-4 % SLOWER on the B200 at every occupancy (43.3-43.8 against 41.9 ms, profiles/r2_dpend_step3_variants.txt): kept as an
// opt-in (SB_DPEND_OPTIMISTIC=1) and as a second implementation the host test compares step() with.
template <class K>
SB_HD void step_optimistic(const Scales& q, Pairs& b, double (&y)[4], const K& k, const bool refresh) {
    const double th1 = y[0], th2 = y[1], w1 = y[2], w2 = y[3];
    const double x1 = k.dt * w1, x2 = k.dt * w2, xd = x1 - x2;
    double a1, a2;
    accel(q.kol, b, w1, w2, a1, a2);                                              // f1
    const double a1f = a1, a2f = a2;
    double u1 = fma(k.hdt, a1, w1), u2 = fma(k.hdt, a2, w2);
    Pairs p;
    rotate_half(b.s1, b.c1, x1, p.s1, p.c1);
    rotate_half(b.S, b.C, xd, p.S, p.C);
    accel(q.kol, p, u1, u2, a1, a2);                                              // f2
    double A1 = a1, A2 = a2;
    const double d1 = k.hdt2 * a1f, dd = fma(-k.hdt2, a2f, d1);
    sbtrig::rotate_tiny_k(p.s1, p.c1, d1, k.rot3, p.s1, p.c1);
    sbtrig::rotate_tiny_k(p.S, p.C, dd, k.rot3, p.S, p.C);
    u1 = fma(k.hdt, a1, w1); u2 = fma(k.hdt, a2, w2);
    accel(q.kol, p, u1, u2, a1, a2);                                              // f3
    A1 += a1; A2 += a2;
    const double h41 = k.dt * u1, h4d = fma(-k.dt, u2, h41);
    rotate_2m6(b.s1, b.c1, h41, k.rot3, k.cot4, p.s1, p.c1);
    rotate_2m6(b.S, b.C, h4d, k.rot3, k.cot4, p.S, p.C);
    u1 = fma(k.dt, a1, w1); u2 = fma(k.dt, a2, w2);
    accel(q.kol, p, u1, u2, a1, a2);                                              // f4
    const double g1 = fma(k.dt2_6, a1f + A1, x1), g2 = fma(k.dt2_6, a2f + A2, x2);
    const double n1 = th1 + g1, n2 = th2 + g2;
    const double gd = g1 - g2;
    const double e1 = g1 - h41, ed = gd - h4d;
    const int m6 = imax(imax(hi_abs(x1), hi_abs(xd)), imax(hi_abs(h41), hi_abs(h4d)));
    const int mt = imax(hi_abs(d1), hi_abs(dd)), mm = imax(hi_abs(e1), hi_abs(ed));
    if (m6 <= 0x3f900000 && mt <= 0x3f200000 && mm <= 0x3ed00000) {
        y[0] = n1; y[1] = n2;
        y[2] = fma(k.dt6, fma(2.0, A1, a1f + a1), w1);
        y[3] = fma(k.dt6, fma(2.0, A2, a2f + a2), w2);
        if (!refresh) {
            sbtrig::rotate_micro(p.s1, p.c1, e1, b.s1, b.c1);
            sbtrig::rotate_micro(p.S, p.C, ed, b.S, b.C);
        } else {
            pairs_of<true>(q, n1, n1 - n2, b);
        }
    } else {
        double ys[4] = {th1, th2, w1, w2};                                        // copies: only they have their address taken
        Pairs bs = b;
        step_out_of_line(q, bs, ys, k, refresh);
        y[0] = ys[0]; y[1] = ys[1]; y[2] = ys[2]; y[3] = ys[3];
        b = bs;
    }
}

}  // namespace sbdp
