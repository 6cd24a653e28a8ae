// trig.cuh -- lean double-precision sin / sincos for the FP64-bound ensemble kernels.
//
// Why: in rk4_ensemble_kernel<DpendSys> built on CUDA's libdevice sin()/sincos(), only 39 % of the
// executed instructions were FP64 (ncu, profiles/r1_dpend_opmix_libdevice.txt): per instance-step 178
// UMOV (64-bit polynomial constants re-materialised in the loop), 64 FSEL, 56 LOP3, 52 MOV, 24 LDG of
// the coefficient table, 24 F2I/I2F and 32 BSSY/BSYNC for the Payne-Hanek slow path kept the issue
// slots 77 % busy while the FP64 pipe sat at 60 %.  The routines below keep the same accuracy class
// (Cody-Waite 3-term reduction with FMA, fdlibm __kernel_sin/__kernel_cos minimax polynomials,
// error ~1 ulp on the reduced range) with ~10 non-FP64 instructions per call:
//   * quadrant by the 1.5*2^52 magic-number add (no F2I/I2F, which run on the quarter-rate XU pipe),
//   * coefficients in uniform registers for the whole time loop (see c_trig below),
//   * no data-dependent branch and no slow path: |x| <= 1e9 is the caller's contract (checked once per
//     derivative call; beyond it the instance is flagged SOPOT_B200_ST_RANGE).
// Only the FP_CONTRACT instantiations use these; FP_STRICT keeps libdevice.
#pragma once

#ifdef __CUDACC__
#define SB_HD __host__ __device__ __forceinline__
#else
#define SB_HD inline
#include <cmath>
#include <cstdint>
#include <cstring>
#endif

namespace sbtrig {

constexpr double kTwoOverPi = 0x1.45f306dc9c883p-1;
constexpr double kMagic = 6755399441055744.0;            // 1.5 * 2^52
constexpr double kPio2Hi = 0x1.921fb54442d18p+0;         // pi/2 = Hi + Mid + Lo to ~160 bits
constexpr double kPio2Mid = 0x1.1a62633145c07p-54;
constexpr double kPio2Lo = -0x1.f1976b7ed8fbcp-110;
constexpr double kMaxArg = 1.0e9;                        // measured: <= 1.9 ulp for |x| <= 1e6 (tests/cpp/trig_accuracy.cpp)
// fdlibm k_sin.c / k_cos.c minimax coefficients on [-pi/4, pi/4]
constexpr double kS1 = -1.66666666666666324348e-01, kS2 = 8.33333333332248946124e-03,
                 kS3 = -1.98412698298579493134e-04, kS4 = 2.75573137070700676789e-06,
                 kS5 = -2.50507602534068634195e-08, kS6 = 1.58969099521155010221e-10;
constexpr double kC1 = 4.16666666666666019037e-02, kC2 = -1.38888888888741095749e-03,
                 kC3 = 2.48015872894767294178e-05, kC4 = -2.75573143513906633035e-07,
                 kC5 = 2.08757232129817482790e-09, kC6 = -1.13596475577881948265e-11;

// Taylor coefficients of sin h / cos h for the small-increment rotation (|h| <= kSmallAngle)
constexpr double kT3 = -1.0 / 6.0, kT5 = 1.0 / 120.0, kT7 = -1.0 / 5040.0;
constexpr double kU2 = -0.5, kU4 = 1.0 / 24.0, kU6 = -1.0 / 720.0;
constexpr double kSmallAngle = 0.03125;                  // 2^-5

SB_HD int lo_word(double x) {
#ifdef __CUDA_ARCH__
    return __double2loint(x);
#else
    uint64_t u; std::memcpy(&u, &x, 8); return (int)(uint32_t)u;
#endif
}
SB_HD double flip_sign_if(double x, int neg_mask_bit31) {     // xor bit 31 of the high word
#ifdef __CUDA_ARCH__
    return __hiloint2double(__double2hiint(x) ^ neg_mask_bit31, __double2loint(x));
#else
    uint64_t u; std::memcpy(&u, &x, 8); u ^= (uint64_t)(uint32_t)neg_mask_bit31 << 32; std::memcpy(&x, &u, 8); return x;
#endif
}

// Coefficient placement matters more than instruction count on sm_100a (tools/microbench/fp64_mix.cu,
// profiles/r1_dpend_notes.md): a DFMA whose three sources are three distinct vector registers is limited
// by register-file read bandwidth to ~3 cycles per warp instruction, 2 cycles with an operand from a
// UNIFORM register.  Literals are re-materialised by ptxas with an UMOV pair at every use, per-thread
// registers cost the third RF read, but a __constant__ table read with STATIC indices is loaded once
// (LDCU.128) into uniform registers that stay live across the whole time loop and feed DFMA directly.
#ifdef __CUDACC__
__constant__ double c_trig[23] = {kTwoOverPi, kMagic, kPio2Hi, kPio2Mid, kPio2Lo, kS1, kS2, kS3, kS4, kS5, kS6,
                                  kC1, kC2, kC3, kC4, kC5, kC6, kT3, kT5, kT7, kU2, kU4, kU6};
#endif
constexpr double h_trig[23] = {kTwoOverPi, kMagic, kPio2Hi, kPio2Mid, kPio2Lo, kS1, kS2, kS3, kS4, kS5, kS6,
                               kC1, kC2, kC3, kC4, kC5, kC6, kT3, kT5, kT7, kU2, kU4, kU6};
#ifdef __CUDA_ARCH__
#define SB_TRIG(i) c_trig[i]
#else
#define SB_TRIG(i) h_trig[i]
#endif

// COLD = true: for a call site on a rarely taken path inside a hot loop.  The 17 constants of the full routine are then
// fetched where they are used (a volatile constant-bank load that the compiler cannot hoist) instead of being kept in
// uniform registers across the loop: with them resident (34 of the 63 uniform registers) ptxas moved the loop's own
// coefficients and step constants to vector registers, which turns their FMAs into three-register ones (3 issue cycles).
template <bool COLD>
SB_HD double trig_const(int i) {
#ifdef __CUDA_ARCH__
    if constexpr (COLD) {
        double v;
        asm volatile("{ .reg .u64 a; cvta.to.const.u64 a, %1; ld.const.f64 %0, [a]; }" : "=d"(v) : "l"(&c_trig[i]));
        return v;
    } else {
        return c_trig[i];
    }
#else
    return h_trig[i];
#endif
}

// x = q*(pi/2) + r, |r| <= pi/4 (+ rounding), q returned modulo 2^32 (only its low 2 bits matter)
template <bool COLD = false>
SB_HD void reduce(double x, double& r, int& q) {
    const double magic = trig_const<COLD>(1);
    double j = fma(x, trig_const<COLD>(0), magic);
    q = lo_word(j);
    j -= magic;
    r = fma(-j, trig_const<COLD>(2), x);
    r = fma(-j, trig_const<COLD>(3), r);
    r = fma(-j, trig_const<COLD>(4), r);
}

// both sin and cos of x (two polynomials, quadrant swap and sign fix)
template <bool COLD = false>
SB_HD void sincos(double x, double& s, double& c) {
    double r; int q;
    reduce<COLD>(x, r, q);
    const double z = r * r;
    double ps = fma(z, trig_const<COLD>(10), trig_const<COLD>(9)); ps = fma(z, ps, trig_const<COLD>(8)); ps = fma(z, ps, trig_const<COLD>(7));
    ps = fma(z, ps, trig_const<COLD>(6)); ps = fma(z, ps, trig_const<COLD>(5));
    double pc = fma(z, trig_const<COLD>(16), trig_const<COLD>(15)); pc = fma(z, pc, trig_const<COLD>(14)); pc = fma(z, pc, trig_const<COLD>(13));
    pc = fma(z, pc, trig_const<COLD>(12)); pc = fma(z, pc, trig_const<COLD>(11));
    // sin(r) = r (1 + z ps), cos(r) = 1 + z (-1/2 + z pc): Horner forms whose FMAs read at most two distinct
    // vector registers (2-cycle issue; the fdlibm forms fma(z*r, ps, r) / fma(z*z, pc, 1 - z/2) need 3-register
    // FMAs at 3 cycles).  Measured <= 1.7 ulp (tests/test_oracle.py::test_lean_trig_accuracy).
    const double sr = r * fma(z, ps, 1.0);
    const double cr = fma(z, fma(z, pc, -0.5), 1.0);
    const bool odd = q & 1;
    const double sv = odd ? cr : sr, cv = odd ? sr : cr;
    s = flip_sign_if(sv, (q << 30) & 0x80000000);           // sin < 0 in quadrants 2, 3
    c = flip_sign_if(cv, ((q + 1) << 30) & 0x80000000);     // cos < 0 in quadrants 1, 2
}

// sin/cos of (x + h) from s = sin x, c = cos x and a small increment h, |h| <= kSmallAngle = 2^-5:
//   sin h = h + h^3 (T3 + h^2 (T5 + h^2 T7))          truncation h^9/9!  <= 7.8e-20
//   cos h - 1 = h^2 (U2 + h^2 (U4 + h^2 U6))          truncation h^8/8!  <= 2.2e-17
//   sin(x+h) = s + (s (cos h - 1) + c sin h),  cos(x+h) = c + (c (cos h - 1) - s sin h)
// 12 FP64 instructions against 21 for a full sincos; absolute error <= 2 ulp(1).  Used by the RK4 stages 2-4,
// whose arguments are the step's base angle plus dt/2 or dt times an angular rate.
SB_HD bool small_angle(double h) {
#ifdef __CUDA_ARCH__
    return (__double2hiint(h) & 0x7fffffff) <= 0x3fa00000;    // |h| <= 2^-5, integer compare (ALU pipe, not FP64)
#else
    return std::fabs(h) <= kSmallAngle;
#endif
}
SB_HD void rotate_small(double s, double c, double h, double& s_out, double& c_out) {
    // Operand placement (tools/microbench/fp64_operands.cu): a DFMA reading three distinct vector registers
    // issues every 3 cycles, any FP64 instruction with <= 2 distinct vector sources every 2.  The forms below keep
    // the three-register FMAs to the two that the rotation itself needs.
    const double h2 = h * h;
    double p = fma(h2, SB_TRIG(19), SB_TRIG(18)); p = fma(h2, p, SB_TRIG(17));
    const double sh = h * fma(h2, p, 1.0);                  // sin h
    double q = fma(h2, SB_TRIG(22), SB_TRIG(21)); q = fma(h2, q, SB_TRIG(20));
    const double ch = fma(h2, q, 1.0);                      // cos h
    s_out = fma(s, ch, c * sh);
    c_out = fma(c, ch, -(s * sh));
}

// The same for a tiny increment, |d| <= kTinyAngle = 2^-13:  sin d = d (1 + T3 d^2) (next term d^5/120: relative 2^-52/120),
// cos d = 1 - d^2/2 (next term d^4/24 <= 9.3e-18 absolute): 8 FP64 instructions.  RK4 stage 3 differs from stage 2
// by (dt/2)^2 times an acceleration, which is such an increment for any sensible step.
constexpr double kTinyAngle = 0x1p-13;
SB_HD bool tiny_angle(double d) {
#ifdef __CUDA_ARCH__
    return (__double2hiint(d) & 0x7fffffff) <= 0x3f200000;    // |d| <= 2^-13
#else
    return std::fabs(d) <= kTinyAngle;
#endif
}
SB_HD void rotate_tiny(double s, double c, double d, double& s_out, double& c_out) {
    const double d2 = d * d;
    const double sd = d * fma(d2, SB_TRIG(17), 1.0);        // sin d
    const double cd = fma(d2, SB_TRIG(20), 1.0);            // cos d
    s_out = fma(s, cd, c * sd);
    c_out = fma(c, cd, -(s * sd));
}

// The two rotations with the coefficients as arguments (for a caller that holds T3, T5, U4 in kernel parameters, i.e.
// constant-bank operands).  T7, U6 truncated to their high word (relative 6e-8 on terms <= 2^-42 of the result) and
// U2 = -1/2 are instruction immediates, so no FMA of the two polynomials reads more than two vector registers.
constexpr double kT7hi = -0x1.a01a000000000p-13, kU6hi = -0x1.6c16c00000000p-10;
SB_HD void rotate_small_k(double s, double c, double h, double T3, double T5, double U4, double& s_out, double& c_out) {
    const double h2 = h * h;
    double p = fma(h2, kT7hi, T5); p = fma(h2, p, T3);
    const double sh = h * fma(h2, p, 1.0);
    double q = fma(h2, kU6hi, U4); q = fma(h2, q, -0.5);
    const double ch = fma(h2, q, 1.0);
    s_out = fma(s, ch, c * sh);
    c_out = fma(c, ch, -(s * sh));
}
SB_HD void rotate_tiny_k(double s, double c, double d, double T3, double& s_out, double& c_out) {
    const double d2 = d * d;
    const double sd = d * fma(d2, T3, 1.0);
    const double cd = fma(d2, -0.5, 1.0);
    s_out = fma(s, cd, c * sd);
    c_out = fma(c, cd, -(s * sd));
}

// The same for a micro increment, |d| <= 2^-18:  sin d = d (next term d^3/6 <= 2^-56.6 of the unit pair), cos d = 1 - d^2/2:
//   s' = s + d (c - (d/2) s),  c' = c - d (s + (d/2) c)        5 FP64 instructions, 14 issue cycles against 18 for rotate_tiny.
// The step from the stage-4 angle to the new base angle of an RK4 step, dt^2/6 (a1 - 2 a2 + a3) = O(dt^3), is always such an increment.
constexpr double kMicroAngle = 0x1p-18;
SB_HD bool micro_angle(double d) {
#ifdef __CUDA_ARCH__
    return (__double2hiint(d) & 0x7fffffff) <= 0x3ed00000;    // |d| <= 2^-18
#else
    return std::fabs(d) <= kMicroAngle;
#endif
}
// -d/2 for the micro rotation.  SB_MICRO_INT_HALF (opt-in, measured 1 % SLOWER on the double pendulum: 42.36 against 41.93 ms per
// 1000 steps -- two FP64 instructions fewer, but SEL / LOP3 / IADD on the step's loop-carried chain, profiles/r2_dpend_step3_variants.txt):
// on the integer pipe (exponent field minus one, sign flipped) instead of a DMUL on the FP64 pipe the kernel is bound by; |d| below 2^-1021 (exponent field < 2, zero included) gives 0, which leaves the
// rotation's second-order term out exactly where it underflows anyway.  Callers guarantee |d| <= 2^-18 (no inf / NaN).
SB_HD double neg_half(double d) {
#if defined(__CUDA_ARCH__) && defined(SB_MICRO_INT_HALF)
    const int hi = __double2hiint(d);
    const bool normal = (hi & 0x7fe00000) != 0;
    return __hiloint2double(normal ? ((hi - 0x00100000) ^ (int)0x80000000) : 0, normal ? __double2loint(d) : 0);
#else
    return -0.5 * d;
#endif
}
SB_HD void rotate_micro(double s, double c, double d, double& s_out, double& c_out) {
    const double eh = neg_half(d);
    const double so = fma(d, fma(eh, s, c), s);
    const double co = fma(d, fma(eh, c, -s), c);
    s_out = so; c_out = co;
}

// 1/x to <= 1 ulp for normal, finite x: MUFU.RCP64H seed (relative error 2^-20, measured:
// tools/microbench/rcp_accuracy.cu) + one cubically convergent step, y (1 + e + e^2) with e = 1 - x y  ->  2^-60.
// No special-case branch (the IEEE division sequence carries a data-dependent slow path; callers guarantee |x|
// in the normal range).
SB_HD double rcp_fast(double x) {
#ifdef __CUDA_ARCH__
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    e = fma(e, e, e);
    return fma(y, e, y);
#else
    return 1.0 / x;
#endif
}

// sqrt(x) and 1/sqrt(x) together for normal, positive x: MUFU.RSQ64H seed (2^-20) + two Newton steps on the
// reciprocal root, one residual correction on the root; both <= 1 ulp, no special-case branch.  Replaces an IEEE
// sqrt (slow-path branch) followed by IEEE divisions by the root.
SB_HD void sqrt_rsqrt_fast(double x, double& root, double& inv_root) {
#ifdef __CUDA_ARCH__
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    double e = fma(-(hx * y), y, 0.5);        // (1 - x y^2) / 2
    y = fma(y, e, y);
    e = fma(-(hx * y), y, 0.5);
    y = fma(y, e, y);
    double r = x * y;
    r = fma(fma(-r, r, x), 0.5 * y, r);
    root = r; inv_root = y;
#else
    root = std::sqrt(x); inv_root = 1.0 / root;
#endif
}

// ---- lean log / exp for the rocket's atmosphere power law  P = P_b (T_b / T)^e = P_b exp(e log(T_b / T)) -------------
// log(num / den) = 2 atanh(z), z = (num - den) / (num + den): no argument reduction, no table, no special cases; the odd
// series z (1 + z^2/3 + ... + z^22/23) truncates below 2e-18 relative for |z| <= kLogRatioMaxZ = 0.18 (0.69 <= num/den
// <= 1.44; inside one layer of the standard atmosphere T_b / T stays within 0.85 .. 1.33).  21 FP64 instructions + MUFU.
// exp(x): k = rint(x / ln 2) by the magic-number add, r = x - k ln2 (hi + lo), Taylor polynomial of degree 13 on
// |r| <= 0.347 (truncation 4e-18), scaling by adding k to the exponent field: valid for |x| <= kExpLeanMax = 690.
// Measured <= 2.4 ulp (log: the quotient z carries the roundings of the sum, the reciprocal and the product) and <= 1 ulp
// (exp), tests/cpp/trig_accuracy.cpp; callers check the ranges and fall back to libm outside.
constexpr double kLogRatioMaxZ = 0.18, kExpLeanMax = 690.0;
constexpr double kLog2e = 0x1.71547652b82fep+0, kLn2Hi = 0x1.62e42fefa39efp-1, kLn2Lo = 0x1.abc9e3b39803fp-56;
#define SB_LNEXP_TABLE {1.0 / 3, 1.0 / 5, 1.0 / 7, 1.0 / 9, 1.0 / 11, 1.0 / 13, 1.0 / 15, 1.0 / 17, 1.0 / 19, 1.0 / 21, 1.0 / 23, \
                        kLog2e, kMagic, kLn2Hi, kLn2Lo, 1.0 / 2, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320, \
                        1.0 / 362880, 1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0}
#ifdef __CUDACC__
__constant__ double c_lnexp[27] = SB_LNEXP_TABLE;
#endif
constexpr double h_lnexp[27] = SB_LNEXP_TABLE;
#ifdef __CUDA_ARCH__
#define SB_LNEXP(i) c_lnexp[i]
#else
#define SB_LNEXP(i) h_lnexp[i]
#endif

SB_HD double log_ratio_z(double num, double den) { return (num - den) * rcp_fast(num + den); }
SB_HD double log_ratio_lean(double num, double den) {       // caller: fabs(log_ratio_z(num, den)) <= kLogRatioMaxZ
    const double z = log_ratio_z(num, den), z2 = z * z;
    double p = fma(z2, SB_LNEXP(10), SB_LNEXP(9));
    p = fma(z2, p, SB_LNEXP(8)); p = fma(z2, p, SB_LNEXP(7)); p = fma(z2, p, SB_LNEXP(6)); p = fma(z2, p, SB_LNEXP(5));
    p = fma(z2, p, SB_LNEXP(4)); p = fma(z2, p, SB_LNEXP(3)); p = fma(z2, p, SB_LNEXP(2)); p = fma(z2, p, SB_LNEXP(1));
    p = fma(z2, p, SB_LNEXP(0));
    const double tz = z + z;
    return fma(tz, z2 * p, tz);
}
SB_HD double exp_lean(double x) {                           // caller: fabs(x) <= kExpLeanMax
    double j = fma(x, SB_LNEXP(11), SB_LNEXP(12));
    const int k = lo_word(j);
    j -= SB_LNEXP(12);
    double r = fma(-j, SB_LNEXP(13), x);
    r = fma(-j, SB_LNEXP(14), r);
    double q = fma(r, SB_LNEXP(26), SB_LNEXP(25));
    q = fma(r, q, SB_LNEXP(24)); q = fma(r, q, SB_LNEXP(23)); q = fma(r, q, SB_LNEXP(22)); q = fma(r, q, SB_LNEXP(21));
    q = fma(r, q, SB_LNEXP(20)); q = fma(r, q, SB_LNEXP(19)); q = fma(r, q, SB_LNEXP(18)); q = fma(r, q, SB_LNEXP(17));
    q = fma(r, q, SB_LNEXP(16)); q = fma(r, q, SB_LNEXP(15));
    const double er = fma(r * r, q, r) + 1.0;               // e^r in (0.70, 1.42)
#ifdef __CUDA_ARCH__
    return __hiloint2double(__double2hiint(er) + (k << 20), __double2loint(er));
#else
    return std::ldexp(er, k);
#endif
}

}  // namespace sbtrig
