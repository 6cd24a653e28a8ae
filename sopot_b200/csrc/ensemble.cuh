// ensemble.cuh -- K1: the fused four-stage RK4 step of core/solver.hpp:91-132 for an ensemble of
// independent instances, one instance per thread, state and stage vectors register-resident for all
// n_steps (no global traffic inside the time loop except optional strided trajectory snapshots).
//
// RK4 association (core/solver.hpp:95-125):  k_i = dt * f(.)  then
//     y += (k1 + 2.0*k2 + 2.0*k3 + k4) / 6.0      evaluated left to right.
// The running accumulator  S <- k1; S <- S + 2.0*k2; S <- S + 2.0*k3; y <- y + (S + k4)/6.0
// reproduces ((k1 + 2k2) + 2k3) + k4 exactly (2.0*k is exact), so only one stage vector is live.
//
// A System provides:
//   static constexpr int DIM, BLOCK, IPT (instances per thread), MIN_BLOCKS (launch-bounds hint);
//   struct DevParams;                       // kernel argument (ParamRef per parameter, tables, ...)
//   template <bool S> struct Inst;          // per-instance constants held in registers
//   template <bool S> static __device__ void load(const DevParams&, size_t i, Inst<S>&, Real<S>(&y)[DIM]);
//   template <bool S> static __device__ int  rhs(const Inst<S>&, const Real<S>(&y)[DIM], Real<S>(&dy)[DIM]);
//   template <bool S> static __device__ void observe(Inst<S>&, double t, const Real<S>(&y)[DIM]);   // after each step
//   template <bool S> static __device__ int  finish(const DevParams&, const Inst<S>&, size_t i, size_t n, const Real<S>(&y)[DIM]);
#pragma once
#include "common.cuh"

template <class Sys, bool S>
__device__ __forceinline__ int rk4_step(const typename Sys::template Inst<S>& in,
                                        Real<S> (&y)[Sys::DIM], const Real<S> dt) {
    constexpr int D = Sys::DIM;
    Real<S> k[D], acc[D], tmp[D];
    int st = Sys::template rhs<S>(in, y, k);
#pragma unroll
    for (int d = 0; d < D; ++d) { k[d] = k[d] * dt; acc[d] = k[d]; tmp[d] = y[d] + 0.5 * k[d]; }
    st |= Sys::template rhs<S>(in, tmp, k);
#pragma unroll
    for (int d = 0; d < D; ++d) { k[d] = k[d] * dt; acc[d] = acc[d] + 2.0 * k[d]; tmp[d] = y[d] + 0.5 * k[d]; }
    st |= Sys::template rhs<S>(in, tmp, k);
#pragma unroll
    for (int d = 0; d < D; ++d) { k[d] = k[d] * dt; acc[d] = acc[d] + 2.0 * k[d]; tmp[d] = y[d] + k[d]; }
    st |= Sys::template rhs<S>(in, tmp, k);
#pragma unroll
    for (int d = 0; d < D; ++d) { k[d] = k[d] * dt; y[d] = y[d] + r_div_const(acc[d] + k[d], 6.0); }
    return st;
}

// FP_CONTRACT form of the same step.  The stage vectors k_i = dt*f_i are never formed: with
// F = f1 + 2 f2 + 2 f3 + f4 the update is y += (dt/6) F and the stage arguments are y + (dt/2) f, y + dt f, each
// one FMA per state element (7 FP64 instructions per element and step instead of 12).  Same real-number
// formula as core/solver.hpp:95-125, different rounding points: differs from the reference by a few ulp of the
// increment, far inside the 1e-12 relative per step this mode promises (tests/test_gpu_parity.py).
// Computed on the HOST and passed as a kernel parameter: a kernel parameter is read as a constant-bank / uniform-register
// operand, so an FMA with one of these and two thread values reads two vector registers (2 issue cycles).  Computed in
// the kernel (0.5 * dt is a per-thread DMUL) the same values sit in vector registers and every such FMA reads three
// (3 cycles, profiles/r1_fp64_operand_rule.txt).  rot3/rot5/cot4: the full-precision coefficients of the small-angle
// rotation (trig.cuh); its remaining coefficients are exact in the high word and become instruction immediates.
struct StepConsts { double dt, hdt, dt6, hdt2, rot3, rot5, cot4, dt2_6; };
__host__ __device__ __forceinline__ StepConsts make_step_consts(double dt) {
    return StepConsts{dt, 0.5 * dt, dt / 6.0, (0.5 * dt) * (0.5 * dt), -1.0 / 6.0, 1.0 / 120.0, 1.0 / 24.0, dt * dt / 6.0};
}

template <class Sys>
__device__ __forceinline__ int rk4_step_fast(const typename Sys::template Inst<false>& in,
                                             Real<false> (&y)[Sys::DIM], const StepConsts& c) {
    constexpr int D = Sys::DIM;
    Real<false> f[D], F[D], tmp[D];
    int st = Sys::template rhs<false>(in, y, f);
#pragma unroll
    for (int d = 0; d < D; ++d) { F[d] = f[d]; tmp[d].v = fma(c.hdt, f[d].v, y[d].v); }
    st |= Sys::template rhs<false>(in, tmp, f);
#pragma unroll
    for (int d = 0; d < D; ++d) { F[d].v = fma(2.0, f[d].v, F[d].v); tmp[d].v = fma(c.hdt, f[d].v, y[d].v); }
    st |= Sys::template rhs<false>(in, tmp, f);
#pragma unroll
    for (int d = 0; d < D; ++d) { F[d].v = fma(2.0, f[d].v, F[d].v); tmp[d].v = fma(c.dt, f[d].v, y[d].v); }
    st |= Sys::template rhs<false>(in, tmp, f);
#pragma unroll
    for (int d = 0; d < D; ++d) y[d].v = fma(c.dt6, F[d].v + f[d].v, y[d].v);
    return st;
}

// dispatch: FP_STRICT -> reference association; FP_CONTRACT -> the system's own fused step when it has one
// (Sys::FAST_STEP, e.g. DpendSys shares the trigonometry between the four stages), else rk4_step_fast.
template <class Sys, bool S>
__device__ __forceinline__ int rk4_step_any(typename Sys::template Inst<S>& in, Real<S> (&y)[Sys::DIM],
                                            const StepConsts& c, const size_t step) {
    if constexpr (S) return rk4_step<Sys, true>(in, y, Real<true>(c.dt));
    else if constexpr (Sys::FAST_STEP) return Sys::fast_step(in, y, c, step);   // may carry values from step to step in `in`
    else return rk4_step_fast<Sys>(in, y, c);
}

// optional hook: Sys::begin_step<S>(Inst&, y) runs once before every step (e.g. a sampled-data controller that
// evaluates its output from the state at the start of the step and holds it over the four stages)
template <class Sys, class = void> struct sb_has_begin_step : std::false_type {};
template <class Sys> struct sb_has_begin_step<Sys, std::void_t<decltype(Sys::HAS_BEGIN_STEP)>> : std::true_type {};

// optional: Sys::MIN_BLOCKS_CONTRACT overrides Sys::MIN_BLOCKS for the FP_CONTRACT instantiation of the RK4 kernel (a system
// whose contract-mode step needs fewer registers than its strict one)
template <class Sys, class = void> struct sb_min_blocks_contract { static constexpr int value = Sys::MIN_BLOCKS; };
template <class Sys> struct sb_min_blocks_contract<Sys, std::void_t<decltype(Sys::MIN_BLOCKS_CONTRACT)>> { static constexpr int value = Sys::MIN_BLOCKS_CONTRACT; };

// IPT = instances per thread (Sys::IPT): IPT independent instances are advanced in lock-step by one
// thread, which gives the scheduler IPT independent dependency chains per warp (FP64 latency hiding)
// while per-system constants held in registers are shared.  Instance j of a thread is
// (blockIdx.x*IPT + j)*blockDim.x + threadIdx.x, so every load and store stays fully coalesced.
template <class Sys, bool S>
__global__ void __launch_bounds__(Sys::BLOCK, S ? Sys::MIN_BLOCKS : sb_min_blocks_contract<Sys>::value)
rk4_ensemble_kernel(const typename Sys::DevParams P, const size_t n, const double t0, const StepConsts sc,
                    const size_t n_steps, const size_t store_every, double* __restrict__ state_out,
                    double* __restrict__ traj_out, int32_t* __restrict__ status) {
    constexpr int D = Sys::DIM;
    constexpr int IPT = Sys::IPT;
    Sys::template block_setup<S>(P);
    const size_t base = (size_t)blockIdx.x * IPT * blockDim.x + threadIdx.x;
    if (base >= n) return;
    typename Sys::template Inst<S> in[IPT];
    Real<S> y[IPT][D];
    size_t idx[IPT];
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
        const size_t i = base + (size_t)j * blockDim.x;
        idx[j] = i < n ? i : base;           // a ragged tail re-computes instance `base` and discards it
        Sys::template load<S>(P, idx[j], in[j], y[j]);
    }
    const double dt_ = sc.dt;
    double t = t0;                           // accumulated like the reference: t += dt (core/solver.hpp:127)
#pragma unroll
    for (int j = 0; j < IPT; ++j) Sys::template observe<S>(in[j], t, y[j]);   // initial condition is stored (:87-88)
    int st[IPT];
#pragma unroll
    for (int j = 0; j < IPT; ++j) st[j] = 0;
    size_t until_snap = store_every, snap = 0;
    for (size_t step = 0; step < n_steps; ++step) {
        if constexpr (sb_has_begin_step<Sys>::value) {
#pragma unroll
            for (int j = 0; j < IPT; ++j) Sys::template begin_step<S>(in[j], y[j]);
        }
#pragma unroll
        for (int j = 0; j < IPT; ++j) st[j] |= rk4_step_any<Sys, S>(in[j], y[j], sc, step);
        t += dt_;
#pragma unroll
        for (int j = 0; j < IPT; ++j) Sys::template observe<S>(in[j], t, y[j]);
        if (store_every && --until_snap == 0) {
            until_snap = store_every;
            if (traj_out) {
#pragma unroll
                for (int j = 0; j < IPT; ++j) {
                    if (j == 0 || base + (size_t)j * blockDim.x < n) {
#pragma unroll
                        for (int d = 0; d < D; ++d) traj_out[(snap * D + d) * n + idx[j]] = y[j][d].v;
                    }
                }
            }
            ++snap;
        }
    }
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
        if (j > 0 && base + (size_t)j * blockDim.x >= n) continue;
        const size_t i = idx[j];
        bool finite = true;
#pragma unroll
        for (int d = 0; d < D; ++d) { state_out[(size_t)d * n + i] = y[j][d].v; finite &= isfinite(y[j][d].v); }
        const int sj = st[j] | Sys::template finish<S>(P, in[j], i, n, y[j]);
        if (status) status[i] = (sj & SOPOT_B200_ST_SINGULAR) ? SOPOT_B200_ST_SINGULAR
                              : !finite ? SOPOT_B200_ST_NONFINITE
                              : (sj & SOPOT_B200_ST_RANGE) ? SOPOT_B200_ST_RANGE : SOPOT_B200_ST_OK;
    }
}

// single derivative evaluation per instance (computeDerivatives), state/out [DIM][n]
template <class Sys, bool S>
__global__ void __launch_bounds__(Sys::BLOCK)
rhs_ensemble_kernel(const typename Sys::DevParams P, const size_t n, const double* __restrict__ state,
                    double* __restrict__ out) {
    constexpr int D = Sys::DIM;
    Sys::template block_setup<S>(P);
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    typename Sys::template Inst<S> in;
    Real<S> y[D], dy[D];
    Sys::template load<S>(P, i, in, y);
#pragma unroll
    for (int d = 0; d < D; ++d) y[d] = Real<S>(state[(size_t)d * n + i]);
    Sys::template rhs<S>(in, y, dy);
#pragma unroll
    for (int d = 0; d < D; ++d) out[(size_t)d * n + i] = dy[d].v;
}

// ---- host-side helpers shared by the ensemble entry points -------------------------------------
struct EnsembleArgs {
    sopot_b200_ctx* ctx;
    size_t n;
    double t0, dt;
    size_t n_steps;
    const sopot_b200_rk4_opts* opts;
    double* state_out;
    double* traj_out;
    int32_t* status;
};

inline int sb_check_ensemble(const EnsembleArgs& a, const void* params) {
    if (!a.ctx) return SOPOT_B200_EINVAL;
    if (!params || !a.opts || !a.state_out)
        return sb_fail(a.ctx, SOPOT_B200_EINVAL, "params, opts and state_out must be non-NULL");
    if (a.opts->mem > 1 || a.opts->fp_mode > 1) return sb_fail(a.ctx, SOPOT_B200_EINVAL, "bad opts");
    if (!(a.dt > 0.0)) return sb_fail(a.ctx, SOPOT_B200_EINVAL, "dt must be positive");
    if (a.traj_out && a.opts->store_every == 0)
        return sb_fail(a.ctx, SOPOT_B200_EINVAL, "traj_out needs store_every > 0");
    return SOPOT_B200_OK;
}

// stage `count` per-instance parameters (each n doubles, or 1 when its broadcast bit is set)
inline int sb_stage_params(Staging& sg, const double* const* user, int count, size_t n,
                           uint32_t broadcast, ParamRef* refs) {
    for (int j = 0; j < count; ++j) {
        if (!user[j]) return sb_fail(sg.ctx, SOPOT_B200_EINVAL, "parameter pointer %d is NULL", j);
        const bool bc = (broadcast >> j) & 1u;
        if (bc) { refs[j].p = nullptr; refs[j].value = *user[j]; continue; }     // host scalar, by value
        const void* dev = nullptr;
        int rc = sg.in(user[j], n * sizeof(double), &dev);
        if (rc) return rc;
        refs[j].p = static_cast<const double*>(dev);
        refs[j].value = 0.0;
    }
    return SOPOT_B200_OK;
}

inline size_t sb_ensemble_arena_bytes(int n_params, int dim, size_t n, size_t n_steps,
                                      const sopot_b200_rk4_opts* o, bool traj, bool status, size_t extra) {
    size_t snaps = (traj && o->store_every) ? n_steps / o->store_every : 0;
    return (size_t)n_params * sb_align(n * 8) + sb_align((size_t)dim * n * 8) +
           sb_align(snaps * dim * n * 8) + (status ? sb_align(n * 4) : 0) + extra + 4096;
}

template <class Sys>
int sb_launch_ensemble(const EnsembleArgs& a, const typename Sys::DevParams& P, Staging& sg) {
    sopot_b200_ctx* ctx = a.ctx;
    const size_t n = a.n;
    const size_t snaps = (a.traj_out && a.opts->store_every) ? a.n_steps / a.opts->store_every : 0;
    void *d_state = nullptr, *d_traj = nullptr, *d_status = nullptr;
    int rc;
    if ((rc = sg.out(a.state_out, (size_t)Sys::DIM * n * 8, &d_state))) return rc;
    if ((rc = sg.out(a.traj_out, snaps * Sys::DIM * n * 8, &d_traj))) return rc;
    if ((rc = sg.out(a.status, n * 4, &d_status))) return rc;
    if (n) {
        const unsigned grid = sb_grid_for(n, Sys::BLOCK * Sys::IPT);
        if (a.opts->fp_mode == SOPOT_B200_FP_STRICT)
            rk4_ensemble_kernel<Sys, true><<<grid, Sys::BLOCK, 0, ctx->stream>>>(
                P, n, a.t0, make_step_consts(a.dt), a.n_steps, a.opts->store_every, (double*)d_state, (double*)d_traj,
                (int32_t*)d_status);
        else
            rk4_ensemble_kernel<Sys, false><<<grid, Sys::BLOCK, 0, ctx->stream>>>(
                P, n, a.t0, make_step_consts(a.dt), a.n_steps, a.opts->store_every, (double*)d_state, (double*)d_traj,
                (int32_t*)d_status);
        SB_CHECK_LAUNCH(ctx);
    }
    return sg.finish();
}
