// lqr.cu -- K5: batched LQR<4,1>::solve (physics/control/lqr.hpp:260-396), the consumer of the K2 Jacobians
// ("A/B Jacobians for LQR sweep", BASELINE config 3), one operating point per thread, every matrix in registers.
//
// Discretisation (:272-283): Ad = I + A dt + (A dt)^2 / 2, Bd = (I dt + A dt * dt/2) B, Qd = Q dt, P0 = Q.
// Iteration (:289-382):  AdTP = Ad' P;  BdTP = Bd' P;  s = R dt + BdTP Bd;  g = BdTP Ad;  Kd = g / s;
//   P' = (AdTP Ad - Kd' g) + Qd;  symmetrise;  stop when ||P' - P||_F / ||P'||_F < tol or ||P' - P||_F < tol;
//   diverged when ||P'||_F > 1e15 or NaN.  Gain (:376-378): K = (B' P) / R.
// Every sum runs in the reference's loop order starting from 0.0, so the Strict instantiation is bit-identical to
// the reference (the only non-arithmetic operation is the correctly rounded sqrt of the Frobenius norm).
// A warp iterates until its slowest lane has converged; finished lanes idle.
#include "common.cuh"
#include "dual.cuh"

namespace {

constexpr int N = 4;

template <bool S>
struct Mat4 { Real<S> a[N][N]; };

template <bool S>
__device__ __forceinline__ Real<S> frob(const Mat4<S>& M) {          // :140-148
    Real<S> sum(0.0);
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) sum = sum + M.a[i][j] * M.a[i][j];
    return r_sqrt(sum);
}

// one LQR<4,1>::solve; returns status, writes K[4] and (optionally) P
template <bool S>
__device__ __forceinline__ int lqr4_solve(const Mat4<S>& A, const Real<S> (&B)[N], const Mat4<S>& Q, const Real<S> R,
                                          const double dt_, const int max_iter, const double tol, Real<S> (&K)[N],
                                          Mat4<S>& P, int& iters) {
    using T = Real<S>;
    const T dt(dt_);
    Mat4<S> Ad, Qd;
    T Bd[N];
    {
        Mat4<S> Adt, Adt2, Bcoef;
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) Adt.a[i][j] = A.a[i][j] * dt;                       // scale(A, dt)
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) {                                                   // multiply(Adt, Adt)
                T s(0.0);
#pragma unroll
                for (int k = 0; k < N; ++k) s = s + Adt.a[i][k] * Adt.a[k][j];
                Adt2.a[i][j] = s;
            }
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const T id(i == j ? 1.0 : 0.0);
                Ad.a[i][j] = (id + Adt.a[i][j]) + Adt2.a[i][j] * T(0.5);                    // add(add(I, Adt), scale(Adt2, .5))
                Bcoef.a[i][j] = id * dt + Adt.a[i][j] * (dt * T(0.5));                      // add(scale(I, dt), scale(Adt, dt*.5))
                Qd.a[i][j] = Q.a[i][j] * dt;                                                // scale(Q, dt)
                P.a[i][j] = Q.a[i][j] * T(1.0);                                             // scale(Q, 1.0)
            }
#pragma unroll
        for (int i = 0; i < N; ++i) {                                                       // multiplyAB(Bcoef, B)
            T s(0.0);
#pragma unroll
            for (int k = 0; k < N; ++k) s = s + Bcoef.a[i][k] * B[k];
            Bd[i] = s;
        }
    }
    const T Rdt = R * dt;
    int status = -1;                     // -1: still iterating
    iters = 0;
    for (int it = 0; it < max_iter && status < 0; ++it) {
        Mat4<S> AdTP;
        T BdTP[N], g[N];
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) {                                                   // multiply(AdT, P)
                T s(0.0);
#pragma unroll
                for (int k = 0; k < N; ++k) s = s + Ad.a[k][i] * P.a[k][j];
                AdTP.a[i][j] = s;
            }
#pragma unroll
        for (int j = 0; j < N; ++j) {                                                       // multiplyBtP(Bd, P)
            T s(0.0);
#pragma unroll
            for (int k = 0; k < N; ++k) s = s + Bd[k] * P.a[k][j];
            BdTP[j] = s;
        }
        T bpb(0.0);
#pragma unroll
        for (int k = 0; k < N; ++k) bpb = bpb + BdTP[k] * Bd[k];
        const T sden = Rdt + bpb;                                                           // R*dt + Bd'PBd
#pragma unroll
        for (int j = 0; j < N; ++j) {                                                       // Bd'P Ad
            T s(0.0);
#pragma unroll
            for (int k = 0; k < N; ++k) s = s + BdTP[k] * Ad.a[k][j];
            g[j] = s;
        }
        if (fabs(sden.v) < 1e-12) { status = SOPOT_B200_ST_SINGULAR; break; }               // the reference throws (:196-198)
        T Kd[N];
#pragma unroll
        for (int j = 0; j < N; ++j) Kd[j] = g[j] / sden;
        Mat4<S> Pn;
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) {
                T s(0.0);
#pragma unroll
                for (int k = 0; k < N; ++k) s = s + AdTP.a[i][k] * Ad.a[k][j];              // multiply(AdTP, Ad)
                const T ktrk = T(0.0) + Kd[i] * g[j];                                       // KtRK (M = 1: one term added to 0.0)
                Pn.a[i][j] = (s - ktrk) + Qd.a[i][j];
            }
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = i + 1; j < N; ++j) {                                               // symmetrise (:352-358)
                const T avg = T(0.5) * (Pn.a[i][j] + Pn.a[j][i]);
                Pn.a[i][j] = avg; Pn.a[j][i] = avg;
            }
        Mat4<S> diff;
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) diff.a[i][j] = Pn.a[i][j] - P.a[i][j];
        const double change = frob(diff).v, p_norm = frob(Pn).v;
        const double rel = (p_norm > 1e-10) ? (Real<S>(change) / Real<S>(p_norm)).v : change;
        P = Pn;
        iters = it + 1;
        if (rel < tol || change < tol) status = 0;                                          // converged (:369)
        else if (p_norm > 1e15 || change != change) status = -2;                            // diverged (:382)
    }
    if (status == SOPOT_B200_ST_SINGULAR) return status;
    if (status != 0) {                                                                      // :388-397
        const double fn = frob(P).v;
        if (!(fn < 1e10) || fn != fn) return SOPOT_B200_ST_NOT_CONVERGED;
    }
    if (fabs(R.v) < 1e-12) return SOPOT_B200_ST_SINGULAR;
#pragma unroll
    for (int j = 0; j < N; ++j) {                                                           // K = (B'P) / R
        T s(0.0);
#pragma unroll
        for (int k = 0; k < N; ++k) s = s + B[k] * P.a[k][j];
        K[j] = s / R;
    }
    return 0;
}

struct LqrQR { double Q[16]; double R; };

template <bool S>
__global__ void __launch_bounds__(128)
lqr4_kernel(const double* __restrict__ Ain, const double* __restrict__ Bin, const size_t P, const LqrQR qr, const double dt,
            const int max_iter, const double tol, double* __restrict__ K, double* __restrict__ Pric,
            int32_t* __restrict__ iters_out, int32_t* __restrict__ status) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    Mat4<S> A, Q, Pm;
    Real<S> B[N], Kg[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int j = 0; j < N; ++j) { A.a[i][j] = Real<S>(__ldg(Ain + (size_t)(i * N + j) * P + p)); Q.a[i][j] = Real<S>(qr.Q[i * N + j]); }
        B[i] = Real<S>(__ldg(Bin + (size_t)i * P + p));
        Kg[i] = Real<S>(0.0);
    }
    int it = 0;
    const int st = lqr4_solve<S>(A, B, Q, Real<S>(qr.R), dt, max_iter, tol, Kg, Pm, it);
    const double bad = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
    for (int j = 0; j < N; ++j) K[(size_t)j * P + p] = st ? bad : Kg[j].v;
    if (Pric) {
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) Pric[(size_t)(i * N + j) * P + p] = Pm.a[i][j].v;
    }
    if (iters_out) iters_out[p] = it;
    if (status) status[p] = st;
}

}  // namespace

SB_EXPORT int sopot_b200_lqr4_gain(sopot_b200_ctx* ctx, const double* A, const double* B, size_t P, const double* Q16,
                                   double R, double riccati_dt, size_t max_iter, double tol,
                                   const sopot_b200_rk4_opts* opts, double* K, double* Pric, int32_t* iters,
                                   int32_t* status) {
    if (!ctx) return SOPOT_B200_EINVAL;
    if (!A || !B || !Q16 || !opts || !K) return sb_fail(ctx, SOPOT_B200_EINVAL, "NULL argument");
    if (opts->mem > 1 || opts->fp_mode > 1) return sb_fail(ctx, SOPOT_B200_EINVAL, "bad opts");
    if (!(riccati_dt > 0.0) || max_iter > 0x7fffffff) return sb_fail(ctx, SOPOT_B200_EINVAL, "bad riccati_dt / max_iter");
    Staging sg(ctx, opts->mem);
    int rc;
    if (sg.host && (rc = sb_arena_begin(ctx, 2 * sb_align(16 * P * 8) + 2 * sb_align(4 * P * 8) + 2 * sb_align(P * 4) + 4096)))
        return rc;
    const void *d_A, *d_B; void *d_K, *d_P, *d_it, *d_st;
    if ((rc = sg.in(A, 16 * P * 8, &d_A))) return rc;
    if ((rc = sg.in(B, 4 * P * 8, &d_B))) return rc;
    if ((rc = sg.out(K, 4 * P * 8, &d_K))) return rc;
    if ((rc = sg.out(Pric, 16 * P * 8, &d_P))) return rc;
    if ((rc = sg.out(iters, P * 4, &d_it))) return rc;
    if ((rc = sg.out(status, P * 4, &d_st))) return rc;
    LqrQR qr;
    for (int i = 0; i < 16; ++i) qr.Q[i] = Q16[i];       // weights are HOST values shared by the sweep
    qr.R = R;
    if (P) {
        const unsigned grid = sb_grid_for(P, 128);
        if (opts->fp_mode == SOPOT_B200_FP_STRICT)
            lqr4_kernel<true><<<grid, 128, 0, ctx->stream>>>((const double*)d_A, (const double*)d_B, P, qr, riccati_dt, (int)max_iter,
                                                             tol, (double*)d_K, (double*)d_P, (int32_t*)d_it, (int32_t*)d_st);
        else
            lqr4_kernel<false><<<grid, 128, 0, ctx->stream>>>((const double*)d_A, (const double*)d_B, P, qr, riccati_dt, (int)max_iter,
                                                              tol, (double*)d_K, (double*)d_P, (int32_t*)d_it, (int32_t*)d_st);
        SB_CHECK_LAUNCH(ctx);
    }
    return sg.finish();
}
