// ho_dpend.cu -- K1 instantiations for the two small analytic systems:
//   config 1  physics::HarmonicOscillator   (physics/harmonic_oscillator.hpp:75-85)
//   config 2  pendulum::DoublePendulum      (physics/pendulum/double_pendulum.hpp:126-179)
#include "dual.cuh"
#include "ensemble.cuh"
#include "trig.cuh"
#ifndef SB_DPEND_STEP
#define SB_DPEND_STEP 3             // 3: the step of dpend_step.cuh (156 FP64 instructions per instance-step, tools/build_variant.py step3 ho_dpend.cu -DSB_DPEND_STEP=3); 2: the round-2 step below (172)
#endif
#if SB_DPEND_STEP >= 3
#include "dpend_step.cuh"
#endif
#ifndef SB_DPEND_OPTIMISTIC
#define SB_DPEND_OPTIMISTIC 0       // 1: sbdp::step_optimistic (one branch per step, second pass through step() when an increment leaves its tier)
#endif

// ------------------------------------------------------------------------------------------------
// HarmonicOscillator::derivatives: {v, T(-k/m)*x - T(c/m)*v}.  The two quotients are loop
// invariant; hoisting them out of the time loop is bit-identical to dividing on every call.
// ------------------------------------------------------------------------------------------------
struct HoSys {
    static constexpr int DIM = 2;
    static constexpr int BLOCK = 128, IPT = 1, MIN_BLOCKS = 1;
    static constexpr bool FAST_STEP = false;
    struct DevParams { ParamRef m, k, c, x0, v0; };
    template <bool S> struct Inst { Real<S> a, b; };   // a = -k/m, b = c/m
    template <bool S> static __device__ __forceinline__ void block_setup(const DevParams&) {}
    template <bool S>
    static __device__ __forceinline__ void load(const DevParams& P, size_t i, Inst<S>& in, Real<S> (&y)[2]) {
        const double m = P.m.at(i), k = P.k.at(i), c = P.c.at(i);
        in.a = Real<S>(-k / m);
        in.b = Real<S>(c / m);
        y[0] = Real<S>(P.x0.at(i));
        y[1] = Real<S>(P.v0.at(i));
    }
    template <bool S>
    static __device__ __forceinline__ int rhs(const Inst<S>& in, const Real<S> (&y)[2], Real<S> (&dy)[2]) {
        dy[0] = y[1];
        dy[1] = in.a * y[0] - in.b * y[1];
        return 0;
    }
    template <bool S> static __device__ __forceinline__ void observe(Inst<S>&, double, const Real<S> (&)[2]) {}
    template <bool S>
    static __device__ __forceinline__ int finish(const DevParams&, const Inst<S>&, size_t, size_t, const Real<S> (&)[2]) { return 0; }
};

SB_EXPORT int sopot_b200_ho_rk4(sopot_b200_ctx* ctx, const sopot_b200_ho_params* p, size_t n, double t0,
                                double dt, size_t n_steps, const sopot_b200_rk4_opts* opts,
                                double* state_out, double* traj_out, int32_t* status) {
    EnsembleArgs a{ctx, n, t0, dt, n_steps, opts, state_out, traj_out, status};
    int rc = sb_check_ensemble(a, p);
    if (rc) return rc;
    Staging sg(ctx, opts->mem);
    if (sg.host && (rc = sb_arena_begin(ctx, sb_ensemble_arena_bytes(5, 2, n, n_steps, opts, traj_out, status, 0))))
        return rc;
    const double* user[5] = {p->m, p->k, p->c, p->x0, p->v0};
    ParamRef r[5];
    if ((rc = sb_stage_params(sg, user, 5, n, opts->broadcast, r))) return rc;
    HoSys::DevParams P{r[0], r[1], r[2], r[3], r[4]};
    return sb_launch_ensemble<HoSys>(a, P, sg);
}

// one derivative evaluation per instance for any ensemble System (computeDerivatives, N instances)
template <class Sys>
static int launch_rhs(sopot_b200_ctx* ctx, const typename Sys::DevParams& P, Staging& sg, size_t n, bool strict,
                      const double* state, double* out) {
    const void* d_state; void* d_out; int rc;
    if ((rc = sg.in(state, Sys::DIM * n * 8, &d_state))) return rc;
    if ((rc = sg.out(out, Sys::DIM * n * 8, &d_out))) return rc;
    if (n) {
        const unsigned grid = sb_grid_for(n, Sys::BLOCK);
        if (strict) rhs_ensemble_kernel<Sys, true><<<grid, Sys::BLOCK, 0, ctx->stream>>>(P, n, (const double*)d_state, (double*)d_out);
        else rhs_ensemble_kernel<Sys, false><<<grid, Sys::BLOCK, 0, ctx->stream>>>(P, n, (const double*)d_state, (double*)d_out);
        SB_CHECK_LAUNCH(ctx);
    }
    return sg.finish();
}

SB_EXPORT int sopot_b200_ho_rhs(sopot_b200_ctx* ctx, const sopot_b200_ho_params* p, size_t n,
                                const sopot_b200_rk4_opts* opts, const double* state, double* out) {
    if (!ctx) return SOPOT_B200_EINVAL;
    if (!p || !opts || !state || !out) return sb_fail(ctx, SOPOT_B200_EINVAL, "NULL argument");
    if (opts->mem > 1 || opts->fp_mode > 1) return sb_fail(ctx, SOPOT_B200_EINVAL, "bad opts: mem and fp_mode must be 0 or 1");
    Staging sg(ctx, opts->mem);
    int rc;
    if (sg.host && (rc = sb_arena_begin(ctx, 5 * sb_align(n * 8) + 2 * sb_align(2 * n * 8) + 4096))) return rc;
    sopot_b200_ho_params q = *p;
    // the initial-condition slots are not read by a derivative evaluation: a missing one aliases `state` as an ARRAY (its
    // broadcast bit is cleared -- a broadcast slot is dereferenced on the host, and `state` may be a device pointer)
    uint32_t bc = opts->broadcast;
    if (!q.x0) { q.x0 = state; bc &= ~(1u << 3); }
    if (!q.v0) { q.v0 = state; bc &= ~(1u << 4); }
    const double* user[5] = {q.m, q.k, q.c, q.x0, q.v0};
    ParamRef r[5];
    if ((rc = sb_stage_params(sg, user, 5, n, bc, r))) return rc;
    HoSys::DevParams P{r[0], r[1], r[2], r[3], r[4]};
    return launch_rhs<HoSys>(ctx, P, sg, n, opts->fp_mode == SOPOT_B200_FP_STRICT, state, out);
}

// ------------------------------------------------------------------------------------------------
// DoublePendulum::derivatives.  Parameter-only products are hoisted in the reference's own
// association: m2*L2*w2*w2*sd = (((m2*L2)*w2)*w2)*sd,  (m1+m2)*g*s1 = ((m1+m2)*g)*s1,
// -L1*w1*w1*sd = (((-L1)*w1)*w1)*sd, so the Strict instantiation performs exactly the reference's
// operations on every call.  sin(delta) and cos(delta) share one argument reduction (sincos).
// In contract mode the two divisions by det become one reciprocal and two multiplies.
// ------------------------------------------------------------------------------------------------
#ifndef SB_DPEND_BLOCK
#define SB_DPEND_BLOCK 128
#endif
#ifndef SB_DPEND_IPT
#define SB_DPEND_IPT 1
#endif
#ifndef SB_DPEND_MINB
#if SB_DPEND_STEP >= 3
#define SB_DPEND_MINB 6             // 80 registers, six CTAs of four warps per SM: 42.08 -> 41.91 ms per 1000 steps of 4 Mi instances (7: 42.06, 8: 42.74)
#else
#define SB_DPEND_MINB 1
#endif
#endif
#ifndef SB_DPEND_MICRO
#define SB_DPEND_MICRO 1            // 0: the round-2 baseline step (rate sums for the angle rows, tiny rotation to the next base) for A/B runs
#endif
#ifndef SB_DPEND_REFRESH
#define SB_DPEND_REFRESH 16         // full sincos of the carried pairs every this many steps (power of two)
#endif
struct DpendSys {
    static constexpr int DIM = 4;
    static constexpr int BLOCK = SB_DPEND_BLOCK, IPT = SB_DPEND_IPT, MIN_BLOCKS = SB_DPEND_MINB;
    static constexpr bool FAST_STEP = true;
    struct DevParams { ParamRef m1, m2, L1, L2, g, th1, th2, w1, w2; };
    // kol, lam, gL1: row-scaled coefficients of the fast step; s1 .. cd: the scaled sin/cos pairs carried from step to step
    template <bool S> struct Inst { Real<S> a11, m2L2, m12g, negL1, L1, L2, g, kol, lam, gL1; double s1, c1, sd, cd; };
    template <bool S> static __device__ __forceinline__ void block_setup(const DevParams&) {}
    template <bool S>
    static __device__ __forceinline__ void load(const DevParams& P, size_t i, Inst<S>& in, Real<S> (&y)[4]) {
        const Real<S> m1(P.m1.at(i)), m2(P.m2.at(i)), L1(P.L1.at(i)), L2(P.L2.at(i)), g(P.g.at(i));
        in.a11 = (m1 + m2) * L1;        // :162
        in.m2L2 = m2 * L2;
        in.m12g = (m1 + m2) * g;
        in.negL1 = -L1;
        in.L1 = L1; in.L2 = L2; in.g = g;
        y[0] = Real<S>(P.th1.at(i)); y[1] = Real<S>(P.th2.at(i));
        y[2] = Real<S>(P.w1.at(i));  y[3] = Real<S>(P.w2.at(i));
        if constexpr (!S) {
            in.lam = L1 / L2;                                   // lam = L1/L2, kap = m2 L2 / ((m1+m2) L1)
            in.kol = (in.m2L2 / in.a11) / in.lam;               // kap / lam
            in.gL1 = g / L1;
            refresh_trig(in, y[0].v, y[1].v);
        }
    }
    template <bool S>
    static __device__ __forceinline__ int rhs(const Inst<S>& in, const Real<S> (&y)[4], Real<S> (&dy)[4]) {
        const Real<S> theta1 = y[0], theta2 = y[1], omega1 = y[2], omega2 = y[3];
        const Real<S> delta = theta1 - theta2;                       // :145
        Real<S> sin_delta, cos_delta, sin_theta1, sin_theta2;
        int st = 0;
        if constexpr (S) {
            r_sincos(delta, sin_delta, cos_delta);
            sin_theta1 = r_sin(theta1); sin_theta2 = r_sin(theta2);
        } else {
            // trig.cuh (<= 1.6 ulp, branch-free; range checked once in finish()).  sin/cos of delta come
            // from the angle-difference identities instead of a third argument reduction + polynomial
            // pair: 46 instead of 63 FP64 operations per derivative call; absolute error <= 4e-16, i.e.
            // far inside the 1e-12 per-step tolerance of this mode.
            Real<S> c1, c2;
            sbtrig::sincos(theta1.v, sin_theta1.v, c1.v);
            sbtrig::sincos(theta2.v, sin_theta2.v, c2.v);
            sin_delta = sin_theta1 * c2 - c1 * sin_theta2;
            cos_delta = c1 * c2 + sin_theta1 * sin_theta2;
        }
        const Real<S> a12 = in.m2L2 * cos_delta;                     // :163
        const Real<S> a21 = in.L1 * cos_delta;                       // :164
        const Real<S> a22 = in.L2;
        const Real<S> b1 = in.m2L2 * omega2 * omega2 * sin_delta - in.m12g * sin_theta1;   // :168
        const Real<S> b2 = in.negL1 * omega1 * omega1 * sin_delta - in.g * sin_theta2;     // :169
        const Real<S> det = in.a11 * a22 - a12 * a21;                // :172
        const Real<S> n1 = b1 * a22 - b2 * a12;                      // :174
        const Real<S> n2 = in.a11 * b2 - a21 * b1;                   // :175
        dy[0] = omega1;
        dy[1] = omega2;
        if constexpr (S) { dy[2] = n1 / det; dy[3] = n2 / det; }
        else { const Real<S> inv(sbtrig::rcp_fast(det.v)); dy[2] = n1 * inv; dy[3] = n2 * inv; }
        return st;
    }
    // ---- FP_CONTRACT fast step -------------------------------------------------------------------------------
    // The Cramer solve of :162-175 with both equations scaled by their parameter-only leading coefficient (row 1 by
    // 1/a11, row 2 by 1/a22):  with kap = m2 L2 / ((m1+m2) L1), lam = L1 / L2, sd/cd = sin/cos(th1 - th2), P = w2^2, Q = w1^2
    //   B1 = kap P sd - (g/L1) s1,   B2 = -lam Q sd - (g/L2) s2,
    //   a1 = (B1 - kap cd B2) / D,   a2 = (B2 - lam cd B1) / D,    D = 1 - kap lam cd^2  in (0, 1].
    // The step works on two SCALED pairs (a rotation is linear, so a scaled pair rotates like the unit one):
    //   (s1, c1) = (g/L1) (sin th1, cos th1)  and  (S, C) = lam (sd, cd).
    // Then (g/L1) s1 and lam cd, lam sd are free, kap cd = (kap/lam) C, kap sd = (kap/lam) S, and
    //   s1 C - c1 S = (g/L1) lam sin th2 = (g/L2) sin th2            [(g/L1) (L1/L2) = g/L2 exactly]
    // is the second gravity term itself: no third pair, no multiplications by g/L.
    // 16 FP64 instructions + 1 MUFU; 5 of them read three distinct vector registers.
    static __device__ __forceinline__ void accel(const Inst<false>& in, double s1, double c1, double S, double C,
                                                 double w1, double w2, double& a1, double& a2) {
        const double g2 = fma(s1, C, -(c1 * S));                    // (g/L2) sin(th2)
        const double kc = in.kol.v * C;                             // kap cd
        const double B1 = fma(in.kol.v * S, w2 * w2, -s1);
        const double B2 = -fma(S, w1 * w1, g2);
        const double D = fma(-kc, C, 1.0);
        const double inv = sbtrig::rcp_fast(D);
        a1 = fma(-kc, B2, B1) * inv;
        a2 = fma(-C, B1, B2) * inv;
    }
    // full evaluation of the carried values from the angles (start of the run, every SB_DPEND_REFRESH-th step, after a large increment)
    static __device__ __forceinline__ void trig_of(const Inst<false>& in, double th1, double delta, double& s1, double& c1,
                                                   double& S, double& C) {
        double st, ct, sd, cd;
        sbtrig::sincos<true>(th1, st, ct);                          // cold call sites: constants loaded on use
        sbtrig::sincos<true>(delta, sd, cd);
        s1 = in.gL1.v * st; c1 = in.gL1.v * ct;
        S = in.lam.v * sd; C = in.lam.v * cd;
    }
    static __device__ __forceinline__ void refresh_trig(Inst<false>& in, double th1, double th2) {
        trig_of(in, th1, th1 - th2, in.s1, in.c1, in.sd, in.cd);     // delta rounded as in the reference (:145)
    }
    // (g/L1) sin/cos of (th1 + h1) and lam sin/cos of (delta + hd) from the base values: rotation by the small increments,
    // full routine when an increment is large.  (Computing the rotation unconditionally and overriding it in a rarely
    // taken branch measured 3 % slower: ptxas adds WARPSYNC / BREAK / register moves around the override.)
    static __device__ __forceinline__ void stage_trig(const Inst<false>& in, const StepConsts& k, const double th1, const double th2,
                                                      const double s1, const double c1, const double S, const double C,
                                                      const double h1, const double hd,
                                                      double& s1o, double& c1o, double& So, double& Co) {
        if (sbtrig::small_angle(h1) && sbtrig::small_angle(hd)) {
            sbtrig::rotate_small_k(s1, c1, h1, k.rot3, k.rot5, k.cot4, s1o, c1o);
            sbtrig::rotate_small_k(S, C, hd, k.rot3, k.rot5, k.cot4, So, Co);
        } else {
            trig_of(in, th1 + h1, (th1 - th2) + hd, s1o, c1o, So, Co);
        }
    }
    // One RK4 step, y += (dt/6)(f1 + 2 f2 + 2 f3 + f4) as in rk4_step_fast, with the trigonometry shared between the
    // stages and the steps.  Every argument of sin/cos in the step is the base angle plus a small increment:
    //   stage 2: th + (dt/2) w            -> rotate the base pair (sbtrig::rotate_small, |h| <= 2^-5)
    //   stage 3: stage 2 + (dt/2)^2 a_1   -> rotate the stage-2 pair by a tiny increment (rotate_tiny, |d| <= 2^-13)
    //   stage 4: th + dt u_3              -> rotate the base pair
    //   next step: th_new - stage 4       -> = dt^2/6 (a_1 - 2 a_2 + a_3): rotate the stage-4 pair by a micro increment (rotate_micro,
    //                                        |e| <= 2^-18; tiny / small rotation as fall-backs); full sincos every SB_DPEND_REFRESH-th
    //                                        (16th) step bounds the accumulated rounding of the carried pair to a few 1e-16
    // Every shortcut has an integer-compare guard on its increment and falls back to the full routine.
    // ~180 FP64 instructions per instance-step (228 with one full sincos pair per step, ~310 with four).
#if SB_DPEND_STEP >= 3
    // third form: csrc/dpend_step.cuh (host-compilable; tests/cpp/dpend_step_host.cpp checks it against a long-double RK4)
    static __device__ __forceinline__ int fast_step(Inst<false>& in, Real<false> (&y)[4], const StepConsts& k, const size_t step) {
        const sbdp::Scales q{in.kol.v, in.lam.v, in.gL1.v};
        sbdp::Pairs b{in.s1, in.c1, in.sd, in.cd};
        double v[4] = {y[0].v, y[1].v, y[2].v, y[3].v};
#if SB_DPEND_OPTIMISTIC
        sbdp::step_optimistic(q, b, v, k, ((step + 1) & (SB_DPEND_REFRESH - 1)) == 0);
#else
        sbdp::step(q, b, v, k, ((step + 1) & (SB_DPEND_REFRESH - 1)) == 0);
#endif
        y[0].v = v[0]; y[1].v = v[1]; y[2].v = v[2]; y[3].v = v[3];
        in.s1 = b.s1; in.c1 = b.c1; in.sd = b.S; in.cd = b.C;
        return 0;
    }
#else
    static __device__ __forceinline__ int fast_step(Inst<false>& in, Real<false> (&y)[4], const StepConsts& k, const size_t step) {
        const double th1 = y[0].v, th2 = y[1].v, w1 = y[2].v, w2 = y[3].v;
        const double s1 = in.s1, c1 = in.c1, S = in.sd, C = in.cd;
        double a1, a2;
        accel(in, s1, c1, S, C, w1, w2, a1, a2);                                    // f1 = (w1, w2, a1, a2)
        double F0 = w1, F1 = w2, F2 = a1, F3 = a2;
        double A1 = a1, A2 = a2;                                                    // a_1 + a_2 + a_3 per angle (SB_DPEND_MICRO)
        double u1 = fma(k.hdt, a1, w1), u2 = fma(k.hdt, a2, w2);                    // stage-2 rates
        const double d1 = k.hdt2 * a1, dd = k.hdt2 * (a1 - a2);                     // stage 3 - stage 2 angles
        double s1s, c1s, Ss, Cs;
        stage_trig(in, k, th1, th2, s1, c1, S, C, k.hdt * w1, k.hdt * (w1 - w2), s1s, c1s, Ss, Cs);
        accel(in, s1s, c1s, Ss, Cs, u1, u2, a1, a2);                                // f2 = (u1, u2, a1, a2)
#if SB_DPEND_MICRO
        F2 = fma(2.0, a1, F2); F3 = fma(2.0, a2, F3); A1 += a1; A2 += a2;
#else
        F0 = fma(2.0, u1, F0); F1 = fma(2.0, u2, F1); F2 = fma(2.0, a1, F2); F3 = fma(2.0, a2, F3);
#endif
        if (sbtrig::tiny_angle(d1) && sbtrig::tiny_angle(dd)) {
            sbtrig::rotate_tiny_k(s1s, c1s, d1, k.rot3, s1s, c1s);
            sbtrig::rotate_tiny_k(Ss, Cs, dd, k.rot3, Ss, Cs);
        } else {
            stage_trig(in, k, th1, th2, s1, c1, S, C, k.hdt * u1, k.hdt * (u1 - u2), s1s, c1s, Ss, Cs);
        }
        u1 = fma(k.hdt, a1, w1); u2 = fma(k.hdt, a2, w2);                           // stage-3 rates
        accel(in, s1s, c1s, Ss, Cs, u1, u2, a1, a2);                                // f3
#if SB_DPEND_MICRO
        F2 = fma(2.0, a1, F2); F3 = fma(2.0, a2, F3); A1 += a1; A2 += a2;
#else
        F0 = fma(2.0, u1, F0); F1 = fma(2.0, u2, F1); F2 = fma(2.0, a1, F2); F3 = fma(2.0, a2, F3);
#endif
        const double h41 = k.dt * u1, h4d = k.dt * (u1 - u2);
        stage_trig(in, k, th1, th2, s1, c1, S, C, h41, h4d, s1s, c1s, Ss, Cs);
        u1 = fma(k.dt, a1, w1); u2 = fma(k.dt, a2, w2);                             // stage-4 rates
        accel(in, s1s, c1s, Ss, Cs, u1, u2, a1, a2);                                // f4
#if SB_DPEND_MICRO
        // angle increments of the whole step in Nystrom form: g = dt w + dt^2/6 (a_1 + a_2 + a_3) (A1, A2 = the sums, kept while
        // F2, F3 were accumulated) -- two instructions fewer than the rate sums F0, F1, and the increment comes without the
        // cancellation of (th + g) - th
        const double g1 = fma(k.dt2_6, A1, k.dt * w1), g2 = fma(k.dt2_6, A2, k.dt * w2);
        const double n1 = th1 + g1, n2 = th2 + g2;
        y[0].v = n1; y[1].v = n2;
        y[2].v = fma(k.dt6, F2 + a1, w1);  y[3].v = fma(k.dt6, F3 + a2, w2);
        const double gd = g1 - g2;
#else
        const double n1 = fma(k.dt6, F0 + u1, th1), n2 = fma(k.dt6, F1 + u2, th2);
        y[0].v = n1; y[1].v = n2;
        y[2].v = fma(k.dt6, F2 + a1, w1);  y[3].v = fma(k.dt6, F3 + a2, w2);
        const double g1 = n1 - th1, gd = g1 - (n2 - th2);
#endif
        // the new base angle differs from the stage-4 angle by dt (F/6 - u_3) = dt^2/6 (a_1 - 2 a_2 + a_3) = O(dt^3): a micro
        // rotation of the stage-4 pair (tiny / small / full as fall-backs)
        const double e1 = g1 - h41, ed = gd - h4d;
        const bool refresh = ((step + 1) & (SB_DPEND_REFRESH - 1)) == 0;
#if SB_DPEND_MICRO
        if (!refresh && sbtrig::micro_angle(e1) && sbtrig::micro_angle(ed)) {
            sbtrig::rotate_micro(s1s, c1s, e1, in.s1, in.c1);
            sbtrig::rotate_micro(Ss, Cs, ed, in.sd, in.cd);
        } else
#endif
        if (!refresh && sbtrig::tiny_angle(e1) && sbtrig::tiny_angle(ed)) {
            sbtrig::rotate_tiny_k(s1s, c1s, e1, k.rot3, in.s1, in.c1);
            sbtrig::rotate_tiny_k(Ss, Cs, ed, k.rot3, in.sd, in.cd);
        } else if (!refresh && sbtrig::small_angle(g1) && sbtrig::small_angle(gd)) {
            sbtrig::rotate_small_k(s1, c1, g1, k.rot3, k.rot5, k.cot4, in.s1, in.c1);
            sbtrig::rotate_small_k(S, C, gd, k.rot3, k.rot5, k.cot4, in.sd, in.cd);
        } else {
            refresh_trig(in, n1, n2);
        }
        return 0;
    }
#endif
    template <bool S> static __device__ __forceinline__ void observe(Inst<S>&, double, const Real<S> (&)[4]) {}
    template <bool S>
    static __device__ __forceinline__ int finish(const DevParams&, const Inst<S>&, size_t, size_t, const Real<S> (&y)[4]) {
        if constexpr (S) return 0;
        else return fmax(fabs(y[0].v), fabs(y[1].v)) > 0.5 * sbtrig::kMaxArg ? SOPOT_B200_ST_RANGE : 0;
    }
};

static int dpend_stage(sopot_b200_ctx* ctx, Staging& sg, const sopot_b200_dpend_params* p, size_t n,
                       uint32_t broadcast, DpendSys::DevParams& P) {
    const double* user[9] = {p->m1, p->m2, p->L1, p->L2, p->g, p->th1, p->th2, p->w1, p->w2};
    ParamRef r[9];
    int rc = sb_stage_params(sg, user, 9, n, broadcast, r);
    if (rc) return rc;
    (void)ctx;
    P = DpendSys::DevParams{r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8]};
    return SOPOT_B200_OK;
}

SB_EXPORT int sopot_b200_dpend_rk4(sopot_b200_ctx* ctx, const sopot_b200_dpend_params* p, size_t n,
                                   double t0, double dt, size_t n_steps, const sopot_b200_rk4_opts* opts,
                                   double* state_out, double* traj_out, int32_t* status) {
    EnsembleArgs a{ctx, n, t0, dt, n_steps, opts, state_out, traj_out, status};
    int rc = sb_check_ensemble(a, p);
    if (rc) return rc;
    Staging sg(ctx, opts->mem);
    if (sg.host && (rc = sb_arena_begin(ctx, sb_ensemble_arena_bytes(9, 4, n, n_steps, opts, traj_out, status, 0))))
        return rc;
    DpendSys::DevParams P;
    if ((rc = dpend_stage(ctx, sg, p, n, opts->broadcast, P))) return rc;
    return sb_launch_ensemble<DpendSys>(a, P, sg);
}

SB_EXPORT int sopot_b200_dpend_rhs(sopot_b200_ctx* ctx, const sopot_b200_dpend_params* p, size_t n,
                                   const sopot_b200_rk4_opts* opts, const double* state, double* out) {
    if (!ctx) return SOPOT_B200_EINVAL;
    if (!p || !opts || !state || !out) return sb_fail(ctx, SOPOT_B200_EINVAL, "NULL argument");
    if (opts->mem > 1 || opts->fp_mode > 1) return sb_fail(ctx, SOPOT_B200_EINVAL, "bad opts: mem and fp_mode must be 0 or 1");
    Staging sg(ctx, opts->mem);
    int rc;
    // the ICs are overwritten by `state`; reuse th1 as a valid pointer when the caller passes none
    sopot_b200_dpend_params q = *p;
    uint32_t bc = opts->broadcast;                     // aliased slots are arrays, never host scalars (see sopot_b200_ho_rhs)
    if (!q.th1) { q.th1 = state; bc &= ~(1u << 5); }
    if (!q.th2) { q.th2 = state; bc &= ~(1u << 6); }
    if (!q.w1) { q.w1 = state; bc &= ~(1u << 7); }
    if (!q.w2) { q.w2 = state; bc &= ~(1u << 8); }
    if (sg.host && (rc = sb_arena_begin(ctx, 9 * sb_align(n * 8) + 2 * sb_align(4 * n * 8) + 4096))) return rc;
    DpendSys::DevParams P;
    if ((rc = dpend_stage(ctx, sg, &q, n, bc, P))) return rc;
    const void* d_state; void* d_out;
    if ((rc = sg.in(state, 4 * n * 8, &d_state))) return rc;
    if ((rc = sg.out(out, 4 * n * 8, &d_out))) return rc;
    if (n) {
        const unsigned grid = sb_grid_for(n, DpendSys::BLOCK);
        if (opts->fp_mode == SOPOT_B200_FP_STRICT)
            rhs_ensemble_kernel<DpendSys, true><<<grid, DpendSys::BLOCK, 0, ctx->stream>>>(P, n, (const double*)d_state, (double*)d_out);
        else
            rhs_ensemble_kernel<DpendSys, false><<<grid, DpendSys::BLOCK, 0, ctx->stream>>>(P, n, (const double*)d_state, (double*)d_out);
        SB_CHECK_LAUNCH(ctx);
    }
    return sg.finish();
}

// ------------------------------------------------------------------------------------------------
// Batched state Jacobians of the double pendulum by forward-mode AD: computeJacobian (core/typed_component.hpp:486-512)
// on DoublePendulum<Dual<double,4>> (tests/double_pendulum_test.cpp:383-412).  One operating point per thread, the
// derivative of :126-179 evaluated once in DualR<4,S> arithmetic with the reference's association; J[(i*4 + j)*n + p] =
// d f_i / d x_j, SoA like the cart-pole linearization (16 coalesced stores per thread).
// ------------------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ void dpend_derivatives_generic(const double m1_, const double m2_, const double L1_, const double L2_,
                                                          const double g_, const T (&y)[4], T (&dy)[4]) {
    using SC = ScalarOf<T>;
    const T theta1 = y[0], theta2 = y[1], omega1 = y[2], omega2 = y[3];
    const T delta = theta1 - theta2;                                             // :145
    T sin_delta, cos_delta;
    r_sincos(delta, sin_delta, cos_delta);
    const T sin_theta1 = r_sin(theta1), sin_theta2 = r_sin(theta2);
    const T m1 = SC::make(m1_), m2 = SC::make(m2_), L1 = SC::make(L1_), L2 = SC::make(L2_), g = SC::make(g_);
    const T a11 = (m1 + m2) * L1;                                                // :162-165
    const T a12 = m2 * L2 * cos_delta;
    const T a21 = L1 * cos_delta;
    const T a22 = L2;
    const T b1 = m2 * L2 * omega2 * omega2 * sin_delta - (m1 + m2) * g * sin_theta1;       // :168
    const T b2 = SC::make(-L1_) * omega1 * omega1 * sin_delta - g * sin_theta2;            // :169
    const T det = a11 * a22 - a12 * a21;                                         // :172
    dy[0] = omega1; dy[1] = omega2;
    dy[2] = (b1 * a22 - b2 * a12) / det;                                         // :174
    dy[3] = (a11 * b2 - a21 * b1) / det;                                         // :175
}

template <bool S>
__global__ void __launch_bounds__(128)
dpend_jacobian_kernel(const DpendSys::DevParams P, const size_t n, const double* __restrict__ state, double* __restrict__ J) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    using D = DualR<4, S>;
    D y[4], dy[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = D::variable(__ldg(state + (size_t)i * n + p), i);
    dpend_derivatives_generic<D>(P.m1.at(p), P.m2.at(p), P.L1.at(p), P.L2.at(p), P.g.at(p), y, dy);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) J[(size_t)(i * 4 + j) * n + p] = dy[i].d[j].v;
}

SB_EXPORT int sopot_b200_dpend_jacobian(sopot_b200_ctx* ctx, const sopot_b200_dpend_params* p, size_t n,
                                        const sopot_b200_rk4_opts* opts, const double* state, double* J) {
    if (!ctx) return SOPOT_B200_EINVAL;
    if (!p || !opts || !state || !J) return sb_fail(ctx, SOPOT_B200_EINVAL, "NULL argument");
    if (opts->mem > 1 || opts->fp_mode > 1) return sb_fail(ctx, SOPOT_B200_EINVAL, "bad opts: mem and fp_mode must be 0 or 1");
    Staging sg(ctx, opts->mem);
    int rc;
    sopot_b200_dpend_params q = *p;                     // the initial-condition arrays are not read
    uint32_t bc = opts->broadcast;                     // aliased slots are arrays, never host scalars (see sopot_b200_ho_rhs)
    if (!q.th1) { q.th1 = state; bc &= ~(1u << 5); }
    if (!q.th2) { q.th2 = state; bc &= ~(1u << 6); }
    if (!q.w1) { q.w1 = state; bc &= ~(1u << 7); }
    if (!q.w2) { q.w2 = state; bc &= ~(1u << 8); }
    if (sg.host && (rc = sb_arena_begin(ctx, 9 * sb_align(n * 8) + sb_align(4 * n * 8) + sb_align(16 * n * 8) + 4096))) return rc;
    DpendSys::DevParams P;
    if ((rc = dpend_stage(ctx, sg, &q, n, bc, P))) return rc;
    const void* d_state; void* d_J;
    if ((rc = sg.in(state, 4 * n * 8, &d_state))) return rc;
    if ((rc = sg.out(J, 16 * n * 8, &d_J))) return rc;
    if (n) {
        const unsigned grid = sb_grid_for(n, 128);
        if (opts->fp_mode == SOPOT_B200_FP_STRICT)
            dpend_jacobian_kernel<true><<<grid, 128, 0, ctx->stream>>>(P, n, (const double*)d_state, (double*)d_J);
        else
            dpend_jacobian_kernel<false><<<grid, 128, 0, ctx->stream>>>(P, n, (const double*)d_state, (double*)d_J);
        SB_CHECK_LAUNCH(ctx);
    }
    return sg.finish();
}

