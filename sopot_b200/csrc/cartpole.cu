// cartpole.cu -- CartNPendulum plant (physics/control/cart_n_pendulum.hpp:143-221, :319-359) on the
// device, generic over the scalar (Real<S> for integration, DualR<N,S> for linearization), plus
//   K1 instantiation  sopot_b200_cartpole_rk4        (constant control force)
//   K2                sopot_b200_cartpole_linearize  = SystemLinearizer<4,1>::linearize
//                     (core/linearization.hpp:75-114) with Dual<double,5>.
#include "dual.cuh"
#include "ensemble.cuh"
#include "trig.cuh"

// Generic N-link plant.  Arrays are indexed with compile-time constants only (full unrolling; row
// swaps of the partial-pivot search are predicated element swaps), so everything stays in registers.
template <int NL, class T>
struct CartPlant {
    static constexpr int NDOF = NL + 1;
    double mc, m[NL], L[NL], g;

    // returns 0 or SOPOT_B200_ST_SINGULAR; deriv/local layout [x, th_1..th_N, xd, w_1..w_N]
    __device__ __forceinline__ int derivatives(const T (&local)[2 * NDOF], const T& F, T (&deriv)[2 * NDOF]) const {
        using SC = ScalarOf<T>;
        const T gT = SC::make(g);
        T s[NL], c[NL], omega[NL];
#pragma unroll
        for (int i = 0; i < NL; ++i) {                    // :155-161
            r_sincos(local[1 + i], s[i], c[i]);
            omega[i] = local[NDOF + 1 + i];
        }
        double Mtail[NL];                                  // tailMass, :134-138
#pragma unroll
        for (int i = 0; i < NL; ++i) {
            double sum = 0.0;
#pragma unroll
            for (int j = i; j < NL; ++j) sum += m[j];
            Mtail[i] = sum;
        }
        T A[NDOF][NDOF], b[NDOF];
        T total = SC::make(mc);                            // :172-174
#pragma unroll
        for (int i = 0; i < NL; ++i) total = total + SC::make(m[i]);
        A[0][0] = total;
        b[0] = F;                                          // :176
#pragma unroll
        for (int a = 0; a < NL; ++a) {                     // :177-182
            const T coupling = SC::make(Mtail[a]) * SC::make(L[a]) * c[a];
            A[0][1 + a] = coupling;
            A[1 + a][0] = coupling;
            b[0] = b[0] + SC::make(Mtail[a]) * SC::make(L[a]) * s[a] * omega[a] * omega[a];
        }
#pragma unroll
        for (int a = 0; a < NL; ++a) {                     // :185-203
            T rhs_a = SC::make(Mtail[a]) * gT * SC::make(L[a]) * s[a];
#pragma unroll
            for (int bb = 0; bb < NL; ++bb) {
                const int mx = a > bb ? a : bb;
                const T cab = c[a] * c[bb] + s[a] * s[bb];
                A[1 + a][1 + bb] = SC::make(Mtail[mx]) * SC::make(L[a]) * SC::make(L[bb]) * cab;
                if (bb != a) {
                    const T sab = s[a] * c[bb] - c[a] * s[bb];
                    rhs_a = rhs_a - SC::make(Mtail[mx]) * SC::make(L[a]) * SC::make(L[bb]) * sab * omega[bb] * omega[bb];
                }
            }
            b[1 + a] = rhs_a;
        }
        // solveLinearSystem, :319-359: Gaussian elimination, partial pivoting on |value|
        int st = 0;
#pragma unroll
        for (int col = 0; col < NDOF; ++col) {
            int pivot = col;
            double best = fabs(sb_value_of(A[col][col]));
#pragma unroll
            for (int row = col + 1; row < NDOF; ++row) {
                const double mag = fabs(sb_value_of(A[row][col]));
                if (mag > best) { best = mag; pivot = row; }
            }
            if (best < 1e-15) st = SOPOT_B200_ST_SINGULAR;  // the reference throws here
#pragma unroll
            for (int row = col + 1; row < NDOF; ++row) {
                if (row == pivot) {
#pragma unroll
                    for (int j = 0; j < NDOF; ++j) { const T t = A[col][j]; A[col][j] = A[row][j]; A[row][j] = t; }
                    const T t = b[col]; b[col] = b[row]; b[row] = t;
                }
            }
#pragma unroll
            for (int row = col + 1; row < NDOF; ++row) {
                const T factor = A[row][col] / A[col][col];
#pragma unroll
                for (int j = col; j < NDOF; ++j) A[row][j] = A[row][j] - factor * A[col][j];
                b[row] = b[row] - factor * b[col];
            }
        }
        T x[NDOF];
#pragma unroll
        for (int i = NDOF - 1; i >= 0; --i) {
            T sum = b[i];
#pragma unroll
            for (int j = i + 1; j < NDOF; ++j) sum = sum - A[i][j] * x[j];
            x[i] = sum / A[i][i];
        }
#pragma unroll
        for (int i = 0; i < NDOF; ++i) { deriv[i] = local[NDOF + i]; deriv[NDOF + i] = x[i]; }   // :206-219
        return st;
    }
};

// ---- FP_CONTRACT step of the one-link plant ------------------------------------------------------------------
// For N = 1 the 2x2 system of :172-203, [[M, mLc], [mLc, mL^2]] [xdd, thdd] = [F + mL s w^2, m g L s] with M = mc + m,
// has the closed form   xdd = (F + mL s w^2 - m g s c) / D,   thdd = (M g s - c (F + mL s w^2)) / (L D),   D = M - m c^2
// (15 FP64 instructions + 1 MUFU instead of the pivoted elimination with three IEEE divisions), and, as in the double
// pendulum (ho_dpend.cu), every sin/cos argument of a step is the base angle plus a small increment: the pair
// (sin th, cos th) is carried and rotated -- stage 2 and 4 from the base pair (rotate_small), stage 3 relative to stage 2
// and the next base relative to stage 4 (rotate_micro, rotate_tiny as fall-back), full sincos every 16th step and whenever an increment leaves its
// guarded range.  D >= mc, so the reference's singular-pivot throw cannot trigger for mc > 0; instances with
// mc <= 1e-12 take the generic step, which reports it.
struct CartFast {
    double mL, mg, Mg, m, M, invL;      // parameter-only products
    double s, c;                        // carried sin / cos of the pole angle
    int generic;                        // mc too small for the closed form's no-singularity argument
    __device__ __forceinline__ void init(double mc, double m_, double L, double g, double th) {
        M = mc + m_; m = m_; mL = m_ * L; mg = m_ * g; Mg = M * g; invL = 1.0 / L;
        generic = !(mc > 1e-12);
        sbtrig::sincos<true>(th, s, c);
    }
    __device__ __forceinline__ void accel(double sn, double cs, double w, double F, double& xdd, double& thdd) const {
        const double b0 = fma(mL * sn, w * w, F);
        const double D = fma(-(m * cs), cs, M);
        const double inv = sbtrig::rcp_fast(D);
        xdd = fma(-(mg * sn), cs, b0) * inv;
        thdd = fma(Mg, sn, -(cs * b0)) * (inv * invL);
    }
    __device__ __forceinline__ void rotate(double s0, double c0, double h, const StepConsts& k, double base_angle, double& so, double& co) const {
        if (sbtrig::small_angle(h)) sbtrig::rotate_small_k(s0, c0, h, k.rot3, k.rot5, k.cot4, so, co);
        else sbtrig::sincos<true>(base_angle + h, so, co);
    }
    // y = [x, th, xd, w]; F held over the step
    __device__ __forceinline__ void step(Real<false> (&y)[4], const double F, const StepConsts& k, const size_t step_no) {
        const double x = y[0].v, th = y[1].v, xd = y[2].v, w = y[3].v;
        double ax, aw;
        accel(s, c, w, F, ax, aw);
        // Nystrom form of the two position-like rows (as ho_dpend.cu): y += dt v + dt^2/6 (a_1 + a_2 + a_3), with the sums
        // Ax, Aw kept next to the velocity rows' F2, F3 = a_1 + 2 a_2 + 2 a_3
        double F2 = ax, F3 = aw, Ax = ax, Aw = aw;
        double uw = fma(k.hdt, aw, w);
        const double d = k.hdt2 * aw;                                  // stage 3 - stage 2 angle
        double ss, cs;
        rotate(s, c, k.hdt * w, k, th, ss, cs);
        accel(ss, cs, uw, F, ax, aw);
        F2 = fma(2.0, ax, F2); F3 = fma(2.0, aw, F3); Ax += ax; Aw += aw;
        if (sbtrig::tiny_angle(d)) sbtrig::rotate_tiny_k(ss, cs, d, k.rot3, ss, cs);
        else rotate(s, c, k.hdt * uw, k, th, ss, cs);
        uw = fma(k.hdt, aw, w);
        accel(ss, cs, uw, F, ax, aw);
        F2 = fma(2.0, ax, F2); F3 = fma(2.0, aw, F3); Ax += ax; Aw += aw;
        const double h4 = k.dt * uw;
        rotate(s, c, h4, k, th, ss, cs);
        uw = fma(k.dt, aw, w);
        accel(ss, cs, uw, F, ax, aw);
        const double gth = fma(k.dt2_6, Aw, k.dt * w), nth = th + gth;
        y[0].v = x + fma(k.dt2_6, Ax, k.dt * xd); y[1].v = nth;
        y[2].v = fma(k.dt6, F2 + ax, xd); y[3].v = fma(k.dt6, F3 + aw, w);
        const double e = gth - h4;                                     // dt^2/6 (a_1 - 2 a_2 + a_3) = O(dt^3)
        const bool refresh = ((step_no + 1) & 15) == 0;
        if (!refresh && sbtrig::micro_angle(e)) sbtrig::rotate_micro(ss, cs, e, s, c);
        else if (!refresh && sbtrig::tiny_angle(e)) sbtrig::rotate_tiny_k(ss, cs, e, k.rot3, s, c);
        else if (!refresh && sbtrig::small_angle(gth)) sbtrig::rotate_small_k(s, c, gth, k.rot3, k.rot5, k.cot4, s, c);
        else sbtrig::sincos<true>(nth, s, c);
    }
};

// ---- K1: cart-pole (N = 1) integration with a constant force ----------------------------------------
struct CartpoleSys {
    static constexpr int DIM = 4;
    static constexpr int BLOCK = 128, IPT = 1, MIN_BLOCKS = 1;
    static constexpr bool FAST_STEP = true;
    struct DevParams { ParamRef mc, m, L, g, x, th, xd, w, force; };
    template <bool S> struct Inst { CartPlant<1, Real<S>> plant; Real<S> F; CartFast fast; };
    static __device__ __forceinline__ int fast_step(Inst<false>& in, Real<false> (&y)[4], const StepConsts& k, const size_t step) {
        if (in.fast.generic) return rk4_step_fast<CartpoleSys>(in, y, k);
        in.fast.step(y, in.F.v, k, step);
        return 0;
    }
    template <bool S> static __device__ __forceinline__ void block_setup(const DevParams&) {}
    template <bool S>
    static __device__ __forceinline__ void load(const DevParams& P, size_t i, Inst<S>& in, Real<S> (&y)[4]) {
        in.plant.mc = P.mc.at(i); in.plant.m[0] = P.m.at(i); in.plant.L[0] = P.L.at(i); in.plant.g = P.g.at(i);
        in.F = Real<S>(P.force.at(i));
        y[0] = Real<S>(P.x.at(i)); y[1] = Real<S>(P.th.at(i)); y[2] = Real<S>(P.xd.at(i)); y[3] = Real<S>(P.w.at(i));
        if constexpr (!S) in.fast.init(in.plant.mc, in.plant.m[0], in.plant.L[0], in.plant.g, y[1].v);
    }
    template <bool S>
    static __device__ __forceinline__ int rhs(const Inst<S>& in, const Real<S> (&y)[4], Real<S> (&dy)[4]) {
        return in.plant.derivatives(y, in.F, dy);
    }
    template <bool S> static __device__ __forceinline__ void observe(Inst<S>&, double, const Real<S> (&)[4]) {}
    template <bool S>
    static __device__ __forceinline__ int finish(const DevParams&, const Inst<S>&, size_t, size_t, const Real<S> (&)[4]) { return 0; }
};

SB_EXPORT int sopot_b200_cartpole_rk4(sopot_b200_ctx* ctx, const sopot_b200_cartpole_params* p, size_t n,
                                      double t0, double dt, size_t n_steps, const sopot_b200_rk4_opts* opts,
                                      double* state_out, double* traj_out, int32_t* status) {
    EnsembleArgs a{ctx, n, t0, dt, n_steps, opts, state_out, traj_out, status};
    int rc = sb_check_ensemble(a, p);
    if (rc) return rc;
    Staging sg(ctx, opts->mem);
    if (sg.host && (rc = sb_arena_begin(ctx, sb_ensemble_arena_bytes(9, 4, n, n_steps, opts, traj_out, status, 0))))
        return rc;
    const double* user[9] = {p->mc, p->m, p->L, p->g, p->x, p->th, p->xd, p->w, p->force};
    ParamRef r[9];
    if ((rc = sb_stage_params(sg, user, 9, n, opts->broadcast, r))) return rc;
    CartpoleSys::DevParams P{r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8]};
    return sb_launch_ensemble<CartpoleSys>(a, P, sg);
}

SB_EXPORT int sopot_b200_cartpole_rhs(sopot_b200_ctx* ctx, const sopot_b200_cartpole_params* p, size_t n,
                                      const sopot_b200_rk4_opts* opts, const double* state, double* out) {
    if (!ctx) return SOPOT_B200_EINVAL;
    if (!p || !opts || !state || !out) return sb_fail(ctx, SOPOT_B200_EINVAL, "NULL argument");
    if (opts->mem > 1 || opts->fp_mode > 1) return sb_fail(ctx, SOPOT_B200_EINVAL, "bad opts: mem and fp_mode must be 0 or 1");
    Staging sg(ctx, opts->mem);
    int rc;
    if (sg.host && (rc = sb_arena_begin(ctx, 9 * sb_align(n * 8) + 2 * sb_align(4 * n * 8) + 4096))) return rc;
    sopot_b200_cartpole_params q = *p;
    uint32_t bc = opts->broadcast;                     // aliased slots are arrays, never host scalars (see sopot_b200_ho_rhs)
    if (!q.x) { q.x = state; bc &= ~(1u << 4); }
    if (!q.th) { q.th = state; bc &= ~(1u << 5); }
    if (!q.xd) { q.xd = state; bc &= ~(1u << 6); }
    if (!q.w) { q.w = state; bc &= ~(1u << 7); }
    const double* user[9] = {q.mc, q.m, q.L, q.g, q.x, q.th, q.xd, q.w, q.force};
    ParamRef r[9];
    if ((rc = sb_stage_params(sg, user, 9, n, bc, r))) return rc;
    CartpoleSys::DevParams P{r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8]};
    const void* d_state; void* d_out;
    if ((rc = sg.in(state, 4 * n * 8, &d_state))) return rc;
    if ((rc = sg.out(out, 4 * n * 8, &d_out))) return rc;
    if (n) {
        const unsigned grid = sb_grid_for(n, CartpoleSys::BLOCK);
        if (opts->fp_mode == SOPOT_B200_FP_STRICT)
            rhs_ensemble_kernel<CartpoleSys, true><<<grid, CartpoleSys::BLOCK, 0, ctx->stream>>>(P, n, (const double*)d_state, (double*)d_out);
        else
            rhs_ensemble_kernel<CartpoleSys, false><<<grid, CartpoleSys::BLOCK, 0, ctx->stream>>>(P, n, (const double*)d_state, (double*)d_out);
        SB_CHECK_LAUNCH(ctx);
    }
    return sg.finish();
}

// ---- K1 closed loop: state feedback with zero-order hold ----------------------------------------------------
// StateFeedbackController<4,1>::compute (state_feedback_controller.hpp:88-104): u = 0; u -= K[j]*(x[j] - ref[j]) for
// j = 0..3 in order; std::clamp(u, u_min, u_max) when saturation is on.  Evaluated in begin_step from the state at
// the start of the step, held over the four stages.
struct CartpoleFbSys {
    static constexpr int DIM = 4;
    static constexpr int BLOCK = 128, IPT = 1, MIN_BLOCKS = 1;
    static constexpr bool FAST_STEP = true;
    static constexpr bool HAS_BEGIN_STEP = true;
    struct DevParams {
        ParamRef mc, m, L, g, x, th, xd, w;
        const double* K;       // [4][n]
        const double* x_ref;   // [4][n] or nullptr
        double u_min, u_max;
        int saturate;
        size_t n;
        double* stats_out;     // [2][n] or nullptr
    };
    template <bool S> struct Inst {
        CartPlant<1, Real<S>> plant;
        Real<S> F, K[4], ref[4];
        double u_min, u_max, max_theta, max_u;
        int saturate;
        CartFast fast;
    };
    static __device__ __forceinline__ int fast_step(Inst<false>& in, Real<false> (&y)[4], const StepConsts& k, const size_t step) {
        if (in.fast.generic) return rk4_step_fast<CartpoleFbSys>(in, y, k);
        in.fast.step(y, in.F.v, k, step);
        return 0;
    }
    template <bool S> static __device__ __forceinline__ void block_setup(const DevParams&) {}
    template <bool S>
    static __device__ __forceinline__ void load(const DevParams& P, size_t i, Inst<S>& in, Real<S> (&y)[4]) {
        in.plant.mc = P.mc.at(i); in.plant.m[0] = P.m.at(i); in.plant.L[0] = P.L.at(i); in.plant.g = P.g.at(i);
        y[0] = Real<S>(P.x.at(i)); y[1] = Real<S>(P.th.at(i)); y[2] = Real<S>(P.xd.at(i)); y[3] = Real<S>(P.w.at(i));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            in.K[j] = Real<S>(__ldg(P.K + (size_t)j * P.n + i));
            in.ref[j] = Real<S>(P.x_ref ? __ldg(P.x_ref + (size_t)j * P.n + i) : 0.0);
        }
        in.u_min = P.u_min; in.u_max = P.u_max; in.saturate = P.saturate;
        in.F = Real<S>(0.0); in.max_theta = 0.0; in.max_u = 0.0;
        if constexpr (!S) in.fast.init(in.plant.mc, in.plant.m[0], in.plant.L[0], in.plant.g, y[1].v);
    }
    template <bool S>
    static __device__ __forceinline__ void begin_step(Inst<S>& in, const Real<S> (&y)[4]) {
        Real<S> u(0.0);
#pragma unroll
        for (int j = 0; j < 4; ++j) u = u - in.K[j] * (y[j] - in.ref[j]);
        if (in.saturate) u = Real<S>(u.v < in.u_min ? in.u_min : (in.u_max < u.v ? in.u_max : u.v));   // std::clamp
        in.F = u;
        in.max_u = fmax(in.max_u, fabs(u.v));
    }
    template <bool S>
    static __device__ __forceinline__ int rhs(const Inst<S>& in, const Real<S> (&y)[4], Real<S> (&dy)[4]) {
        return in.plant.derivatives(y, in.F, dy);
    }
    template <bool S> static __device__ __forceinline__ void observe(Inst<S>& in, double, const Real<S> (&y)[4]) {
        in.max_theta = fmax(in.max_theta, fabs(y[1].v));
    }
    template <bool S>
    static __device__ __forceinline__ int finish(const DevParams& P, const Inst<S>& in, size_t i, size_t n, const Real<S> (&)[4]) {
        if (P.stats_out) { P.stats_out[i] = in.max_theta; P.stats_out[n + i] = in.max_u; }
        return 0;
    }
};

SB_EXPORT int sopot_b200_cartpole_feedback_rk4(sopot_b200_ctx* ctx, const sopot_b200_cartpole_params* p,
                                               const sopot_b200_feedback* fb, size_t n, double t0, double dt,
                                               size_t n_steps, const sopot_b200_rk4_opts* opts, double* state_out,
                                               double* traj_out, double* stats_out, int32_t* status) {
    EnsembleArgs a{ctx, n, t0, dt, n_steps, opts, state_out, traj_out, status};
    int rc = sb_check_ensemble(a, p);
    if (rc) return rc;
    if (!fb || !fb->K) return sb_fail(ctx, SOPOT_B200_EINVAL, "feedback gain K must be given");
    if (fb->saturate && !(fb->u_min <= fb->u_max)) return sb_fail(ctx, SOPOT_B200_EINVAL, "u_min must not exceed u_max");
    Staging sg(ctx, opts->mem);
    if (sg.host && (rc = sb_arena_begin(ctx, sb_ensemble_arena_bytes(8, 4, n, n_steps, opts, traj_out, status,
                                                                     2 * sb_align(4 * n * 8) + sb_align(2 * n * 8)))))
        return rc;
    const double* user[8] = {p->mc, p->m, p->L, p->g, p->x, p->th, p->xd, p->w};
    ParamRef r[8];
    if ((rc = sb_stage_params(sg, user, 8, n, opts->broadcast, r))) return rc;
    const void *d_K, *d_ref; void* d_stats;
    if ((rc = sg.in(fb->K, 4 * n * 8, &d_K))) return rc;
    if ((rc = sg.in(fb->x_ref, 4 * n * 8, &d_ref))) return rc;
    if ((rc = sg.out(stats_out, 2 * n * 8, &d_stats))) return rc;
    CartpoleFbSys::DevParams P{r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], (const double*)d_K, (const double*)d_ref,
                               fb->u_min, fb->u_max, fb->saturate, n, (double*)d_stats};
    return sb_launch_ensemble<CartpoleFbSys>(a, P, sg);
}

// ---- K2: batched linearization, one operating point per thread ------------------------------------
// Dual<double,5> = 6 doubles per scalar, the whole 2x2 solve stays in registers.  Inputs and the 20
// Jacobian entries are SoA, so every load and store of a warp is one fully coalesced 256 B segment:
// 200 algorithmic bytes per point, HBM-bound.
template <bool S>
__global__ void __launch_bounds__(128)
cartpole_linearize_kernel(ParamRef mc, ParamRef m, ParamRef L, ParamRef g, const double* __restrict__ x0,
                          const double* __restrict__ u0, size_t P, double* __restrict__ A,
                          double* __restrict__ B, int32_t* __restrict__ status) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    using D = DualR<5, S>;
    CartPlant<1, D> plant;
    plant.mc = mc.at(p); plant.m[0] = m.at(p); plant.L[0] = L.at(p); plant.g = g.at(p);
    D xs[4], d[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) xs[i] = D::variable(__ldg(x0 + (size_t)i * P + p), i);   // linearization.hpp:86-89
    const D us = D::variable(__ldg(u0 + p), 4);                                           // :92-95
    const int st = plant.derivatives(xs, us, d);
    const double bad = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) A[(size_t)(i * 4 + j) * P + p] = st ? bad : d[i].d[j].v;   // :102-111
        B[(size_t)i * P + p] = st ? bad : d[i].d[4].v;
    }
    if (status) status[p] = st;
}

SB_EXPORT int sopot_b200_cartpole_linearize(sopot_b200_ctx* ctx, const sopot_b200_cartpole_plant* plant,
                                            const double* x0, const double* u0, size_t P,
                                            const sopot_b200_rk4_opts* opts, double* A, double* B,
                                            int32_t* status) {
    if (!ctx) return SOPOT_B200_EINVAL;
    if (!plant || !x0 || !u0 || !opts || !A || !B) return sb_fail(ctx, SOPOT_B200_EINVAL, "NULL argument");
    if (opts->mem > 1 || opts->fp_mode > 1) return sb_fail(ctx, SOPOT_B200_EINVAL, "bad opts: mem and fp_mode must be 0 or 1");
    Staging sg(ctx, opts->mem);
    int rc;
    if (sg.host && (rc = sb_arena_begin(ctx, 4 * sb_align(P * 8) + sb_align(4 * P * 8) + sb_align(P * 8) +
                                                 sb_align(16 * P * 8) + sb_align(4 * P * 8) + sb_align(P * 4) + 4096)))
        return rc;
    const double* user[4] = {plant->mc, plant->m, plant->L, plant->g};
    ParamRef r[4];
    if ((rc = sb_stage_params(sg, user, 4, P, opts->broadcast, r))) return rc;
    const void *d_x, *d_u; void *d_A, *d_B, *d_st;
    if ((rc = sg.in(x0, 4 * P * 8, &d_x))) return rc;
    if ((rc = sg.in(u0, P * 8, &d_u))) return rc;
    if ((rc = sg.out(A, 16 * P * 8, &d_A))) return rc;
    if ((rc = sg.out(B, 4 * P * 8, &d_B))) return rc;
    if ((rc = sg.out(status, P * 4, &d_st))) return rc;
    if (P) {
        const unsigned grid = sb_grid_for(P, 128);
        if (opts->fp_mode == SOPOT_B200_FP_STRICT)
            cartpole_linearize_kernel<true><<<grid, 128, 0, ctx->stream>>>(r[0], r[1], r[2], r[3], (const double*)d_x,
                (const double*)d_u, P, (double*)d_A, (double*)d_B, (int32_t*)d_st);
        else
            cartpole_linearize_kernel<false><<<grid, 128, 0, ctx->stream>>>(r[0], r[1], r[2], r[3], (const double*)d_x,
                (const double*)d_u, P, (double*)d_A, (double*)d_B, (int32_t*)d_st);
        SB_CHECK_LAUNCH(ctx);
    }
    return sg.finish();
}
