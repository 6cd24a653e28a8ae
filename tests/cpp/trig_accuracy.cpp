// Host-side accuracy check of csrc/trig.cuh (the header compiles for the host with the same fma() sequence):
// max error in ulp of sbtrig::sincos against long double libm over |x| <= 1e6 and near multiples of pi/2, and
// of rotate_small / rotate_small_k (sin x, cos x, h) against sin/cos(x + h) for |h| <= 2^-5 and rotate_tiny / rotate_tiny_k
// for |d| <= 2^-13 and rotate_micro for |d| <= 2^-18 (the variants the double-pendulum step uses).  Pure CPU; run by tests/test_oracle.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../../sopot_b200/csrc/trig.cuh"

static double ulp_err(double got, long double want) {
    const double w = (double)want;
    const double u = std::ldexp(1.0, std::ilogb(w == 0.0 ? 1e-300 : w) - 52);
    return (double)std::fabs(((long double)got - want) / u);
}

int main(int argc, char** argv) {
    const long n = argc > 1 ? std::atol(argv[1]) : 400000;
    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    double worst_s = 0, worst_c = 0, worst_rs = 0, worst_rc = 0, worst_k = 0, worst_t = 0, worst_m = 0;
    for (long i = 0; i < n; ++i) {
        double x;
        switch (i & 3) {
            case 0: x = 3.5 * U(rng); break;
            case 1: x = 1e3 * U(rng); break;
            case 2: x = 1e6 * U(rng); break;
            default: x = std::round(40 * U(rng)) * 1.5707963267948966 + 1e-3 * U(rng); break;
        }
        double s, c;
        sbtrig::sincos(x, s, c);
        const long double sl = sinl((long double)x), cl = cosl((long double)x);
        worst_s = std::fmax(worst_s, ulp_err(s, sl));
        worst_c = std::fmax(worst_c, ulp_err(c, cl));
        const double h = sbtrig::kSmallAngle * U(rng);
        if (!sbtrig::small_angle(h)) { std::printf("small_angle rejects %.17g\n", h); return 1; }
        double rs, rc;
        sbtrig::rotate_small((double)sl, (double)cl, h, rs, rc);
        // absolute error in units of ulp(1) = 2^-52: the rotated pair feeds products of magnitude <= 1
        // (x + h is not representable for large x: the exact pair comes from the addition theorem in long double)
        const long double shl = sinl((long double)h), chl = cosl((long double)h);
        worst_rs = std::fmax(worst_rs, (double)std::fabs((long double)rs - (sl * chl + cl * shl)) / 0x1p-52);
        worst_rc = std::fmax(worst_rc, (double)std::fabs((long double)rc - (cl * chl - sl * shl)) / 0x1p-52);
        // the same rotation with the coefficients as arguments (T7, U6 truncated to instruction immediates)
        sbtrig::rotate_small_k((double)sl, (double)cl, h, -1.0 / 6.0, 1.0 / 120.0, 1.0 / 24.0, rs, rc);
        worst_k = std::fmax(worst_k, (double)std::fabs((long double)rs - (sl * chl + cl * shl)) / 0x1p-52);
        worst_k = std::fmax(worst_k, (double)std::fabs((long double)rc - (cl * chl - sl * shl)) / 0x1p-52);
        // tiny increments
        const double d = sbtrig::kTinyAngle * U(rng);
        if (!sbtrig::tiny_angle(d)) { std::printf("tiny_angle rejects %.17g\n", d); return 1; }
        const long double sdl = sinl((long double)d), cdl = cosl((long double)d);
        double ts, tc, ts2, tc2;
        sbtrig::rotate_tiny((double)sl, (double)cl, d, ts, tc);
        sbtrig::rotate_tiny_k((double)sl, (double)cl, d, -1.0 / 6.0, ts2, tc2);
        if (ts != ts2 || tc != tc2) { std::puts("rotate_tiny and rotate_tiny_k differ"); return 1; }
        worst_t = std::fmax(worst_t, (double)std::fabs((long double)ts - (sl * cdl + cl * sdl)) / 0x1p-52);
        worst_t = std::fmax(worst_t, (double)std::fabs((long double)tc - (cl * cdl - sl * sdl)) / 0x1p-52);
        // micro increments (every 8th sample sits on the bound)
        const double e = (i & 7) ? sbtrig::kMicroAngle * U(rng) : ((i & 8) ? sbtrig::kMicroAngle : -sbtrig::kMicroAngle);
        if (!sbtrig::micro_angle(e)) { std::printf("micro_angle rejects %.17g\n", e); return 1; }
        const long double sel = sinl((long double)e), cel = cosl((long double)e);
        double ms, mc;
        sbtrig::rotate_micro((double)sl, (double)cl, e, ms, mc);
        worst_m = std::fmax(worst_m, (double)std::fabs((long double)ms - (sl * cel + cl * sel)) / 0x1p-52);
        worst_m = std::fmax(worst_m, (double)std::fabs((long double)mc - (cl * cel - sl * sel)) / 0x1p-52);
    }
    // lean log(num/den) and exp (rocket atmosphere): relative error in ulp over the guarded ranges
    double worst_l = 0, worst_e = 0;
    for (long i = 0; i < n; ++i) {
        const double den = 150.0 + 200.0 * (0.5 + 0.5 * U(rng));
        const double ratio = 0.70 + 0.73 * (0.5 + 0.5 * U(rng));
        const double num = den * ratio;
        if (std::fabs(sbtrig::log_ratio_z(num, den)) > sbtrig::kLogRatioMaxZ) { std::printf("log_ratio range %.17g\n", ratio); return 1; }
        const long double wl = logl((long double)num / (long double)den);
        if (std::fabs((double)wl) > 1e-3) worst_l = std::fmax(worst_l, ulp_err(sbtrig::log_ratio_lean(num, den), wl));
        else if (std::fabs(sbtrig::log_ratio_lean(num, den) - (double)wl) > 1e-18) { std::puts("log_ratio near 1"); return 1; }
        const double x = (i & 1 ? 40.0 : 690.0) * U(rng);
        worst_e = std::fmax(worst_e, ulp_err(sbtrig::exp_lean(x), expl((long double)x)));
    }
    std::printf("log_ratio_lean max ulp %.3f; exp_lean max ulp %.3f\n", worst_l, worst_e);
    if (worst_l > 3.0 || worst_e > 2.0) return 3;     // the quotient z carries ~2 ulp (sum, reciprocal, product roundings)
    if (sbtrig::small_angle(0.03126) || !sbtrig::small_angle(-0.03125)) { std::puts("small_angle threshold wrong"); return 1; }
    if (sbtrig::tiny_angle(0x1.0001p-13) || !sbtrig::tiny_angle(-0x1p-13)) { std::puts("tiny_angle threshold wrong"); return 1; }
    if (sbtrig::micro_angle(0x1.0001p-18) || !sbtrig::micro_angle(-0x1p-18)) { std::puts("micro_angle threshold wrong"); return 1; }
    std::printf("sincos max ulp: sin %.3f cos %.3f; rotate_small max abs error / 2^-52: sin %.3f cos %.3f, _k %.3f, tiny %.3f, micro %.3f\n",
                worst_s, worst_c, worst_rs, worst_rc, worst_k, worst_t, worst_m);
    return (worst_s <= 2.0 && worst_c <= 2.0 && worst_rs <= 2.0 && worst_rc <= 2.0 && worst_k <= 2.0 && worst_t <= 2.0 && worst_m <= 2.0) ? 0 : 2;
}
