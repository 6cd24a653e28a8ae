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
#include "trig.cuh"

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

template <bool S>
__device__ __forceinline__ Real<S> frob2(const Mat4<S>& M) {         // squared Frobenius norm, same summation order
    Real<S> sum(0.0);
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) sum = sum + M.a[i][j] * M.a[i][j];
    return sum;
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
        if constexpr (S) {
#pragma unroll
            for (int j = 0; j < N; ++j) Kd[j] = g[j] / sden;
        } else {
            const T inv(sbtrig::rcp_fast(sden.v));                                          // |sden| >= 1e-12 checked above
#pragma unroll
            for (int j = 0; j < N; ++j) Kd[j] = g[j] * inv;
        }
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
        P = Pn;
        iters = it + 1;
        if constexpr (S) {
            const double change = frob(diff).v, p_norm = frob(Pn).v;
            const double rel = (p_norm > 1e-10) ? (Real<S>(change) / Real<S>(p_norm)).v : change;
            if (rel < tol || change < tol) status = 0;                                      // converged (:369)
            else if (p_norm > 1e15 || change != change) status = -2;                        // diverged (:382)
        } else {
            // the same tests on the squared norms: two square roots and a division per iteration are the longest links
            // of the dependent chain.  change / p_norm < tol  <=>  change^2 < tol^2 p_norm^2 (all quantities >= 0)
            const double c2 = frob2(diff).v, p2 = frob2(Pn).v, tol2 = tol * tol;
            const double rel2_lhs = c2, rel2_rhs = (p2 > 1e-20) ? tol2 * p2 : tol2;
            if (rel2_lhs < rel2_rhs || c2 < tol2) status = 0;
            else if (p2 > 1e30 || c2 != c2) status = -2;
        }
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

// ---- LQR<NS,1>::solve for larger plants: one CTA of NS x NS threads per operating point --------------------------
// Thread (i, j) owns element (i, j) of every NS x NS matrix; the matrices live in shared memory.  Each dot product
// runs over k = 0..NS-1 starting from 0.0 and the two Frobenius norms are summed by one thread in row-major order, so
// the Strict instantiation is bit-identical to the reference class for any NS (the reference's own tests use the
// 6-link plant: LQR<14,1>, tests/inverted_pendulum_test.cpp:208-232).
template <int NS, bool S>
__global__ void __launch_bounds__(NS * NS)
lqr_block_kernel(const double* __restrict__ Ain, const double* __restrict__ Bin, const size_t P, const double* __restrict__ Qin,
                 const double R_, const double dt_, const int max_iter, const double tol, double* __restrict__ K,
                 double* __restrict__ Pric, int32_t* __restrict__ iters_out, int32_t* __restrict__ status) {
    using T = Real<S>;
    __shared__ double Ad[NS][NS], Pm[NS][NS], W1[NS][NS], Pn[NS][NS], W2[NS][NS];
    __shared__ double Bd[NS], BdTP[NS], g[NS], Kd[NS], Bv[NS];
    __shared__ double s_sden;
    __shared__ int s_state;          // -1 iterating, 0 converged, -2 diverged, SOPOT_B200_ST_SINGULAR
    const size_t p = blockIdx.x;
    const int i = threadIdx.x / NS, j = threadIdx.x % NS, tid = threadIdx.x;
    const T dt(dt_), R(R_);
    W1[i][j] = (T(Ain[(size_t)(i * NS + j) * P + p]) * dt).v;                  // Adt = scale(A, dt)
    if (j == 0) Bv[i] = Bin[(size_t)i * P + p];
    if (tid == 0) s_state = -1;
    __syncthreads();
    const T q(Qin[i * NS + j]);
    const T Qd = q * dt;                                                       // scale(Q, dt)
    {
        T s2(0.0);
        for (int k = 0; k < NS; ++k) s2 = s2 + T(W1[i][k]) * T(W1[k][j]);      // multiply(Adt, Adt)
        const T id(i == j ? 1.0 : 0.0);
        Ad[i][j] = ((id + T(W1[i][j])) + s2 * T(0.5)).v;
        W2[i][j] = (id * dt + T(W1[i][j]) * (dt * T(0.5))).v;                  // Bcoef
        Pm[i][j] = (q * T(1.0)).v;                                             // P0 = scale(Q, 1.0)
    }
    __syncthreads();
    if (j == 0) {
        T s(0.0);
        for (int k = 0; k < NS; ++k) s = s + T(W2[i][k]) * T(Bv[k]);           // Bd = Bcoef B
        Bd[i] = s.v;
    }
    __syncthreads();
    const T Rdt = R * dt;
    int it = 0;
    for (; it < max_iter; ++it) {
        {
            T s(0.0);
            for (int k = 0; k < NS; ++k) s = s + T(Ad[k][i]) * T(Pm[k][j]);    // AdTP = Ad' P
            W1[i][j] = s.v;
        }
        if (i == 0) {
            T s(0.0);
            for (int k = 0; k < NS; ++k) s = s + T(Bd[k]) * T(Pm[k][j]);       // BdTP = Bd' P
            BdTP[j] = s.v;
        }
        __syncthreads();
        if (i == 0) {
            T s(0.0);
            for (int k = 0; k < NS; ++k) s = s + T(BdTP[k]) * T(Ad[k][j]);     // g = Bd'P Ad
            g[j] = s.v;
        }
        if (tid == NS) {                                                       // (a thread of another row: runs beside g)
            T bpb(0.0);
            for (int k = 0; k < NS; ++k) bpb = bpb + T(BdTP[k]) * T(Bd[k]);
            const T sden = Rdt + bpb;
            s_sden = sden.v;
            if (fabs(sden.v) < 1e-12) s_state = SOPOT_B200_ST_SINGULAR;
        }
        __syncthreads();
        if (s_state == SOPOT_B200_ST_SINGULAR) break;
        if (i == 0) Kd[j] = (T(g[j]) / T(s_sden)).v;
        __syncthreads();
        {
            T s(0.0);
            for (int k = 0; k < NS; ++k) s = s + T(W1[i][k]) * T(Ad[k][j]);    // AdTP Ad
            const T ktrk = T(0.0) + T(Kd[i]) * T(g[j]);
            Pn[i][j] = ((s - ktrk) + Qd).v;
        }
        __syncthreads();
        if (i < j) {                                                           // symmetrise: one thread per pair
            const T avg = T(0.5) * (T(Pn[i][j]) + T(Pn[j][i]));
            Pn[i][j] = avg.v; Pn[j][i] = avg.v;
        }
        __syncthreads();
        W2[i][j] = (T(Pn[i][j]) - T(Pm[i][j])).v;                              // diff
        __syncthreads();
        if (tid == 0 || tid == 32) {                                           // the two norms, each summed in order by one thread
            const double (*M)[NS] = tid == 0 ? W2 : Pn;
            T sum(0.0);
            for (int a = 0; a < NS; ++a)
                for (int b = 0; b < NS; ++b) sum = sum + T(M[a][b]) * T(M[a][b]);
            (tid == 0 ? Bv[0] : Bv[1]) = r_sqrt(sum).v;                        // Bv is free after the set-up
        }
        Pm[i][j] = Pn[i][j];                                                   // m_P = P_new
        __syncthreads();
        if (tid == 0) {
            const double change = Bv[0], p_norm = Bv[1];
            const double rel = (p_norm > 1e-10) ? (T(change) / T(p_norm)).v : change;
            if (rel < tol || change < tol) s_state = 0;
            else if (p_norm > 1e15 || change != change) s_state = -2;
        }
        __syncthreads();
        if (s_state != -1) { ++it; break; }
    }
    // Every thread reads the loop's verdict BEFORE thread 0 may rewrite it, and the barriers below are executed by all
    // threads whatever the verdict (a barrier inside `if (st != 0)` would be divergent when a late warp sees the rewritten
    // value): read -> barrier -> thread 0 decides the fallback -> barrier -> read.
    int st = s_state;
    __syncthreads();
    if (tid == 0 && st != SOPOT_B200_ST_SINGULAR && st != 0) {                 // not converged cleanly (:388-397)
        T sum(0.0);
        for (int a = 0; a < NS; ++a)
            for (int b = 0; b < NS; ++b) sum = sum + T(Pm[a][b]) * T(Pm[a][b]);
        const double fn = r_sqrt(sum).v;
        s_state = (!(fn < 1e10) || fn != fn) ? SOPOT_B200_ST_NOT_CONVERGED : 0;
    }
    __syncthreads();
    st = s_state;
    if (st == 0 && fabs(R_) < 1e-12) st = SOPOT_B200_ST_SINGULAR;
    const double bad = __longlong_as_double(0x7ff8000000000000LL);
    if (i == 0) {
        double kv = bad;
        if (st == 0) {
            T s(0.0);
            for (int k = 0; k < NS; ++k) s = s + T(Bin[(size_t)k * P + p]) * T(Pm[k][j]);   // K = (B'P) / R
            kv = (s / R).v;
        }
        K[(size_t)j * P + p] = kv;
    }
    if (Pric) Pric[(size_t)(i * NS + j) * P + p] = Pm[i][j];
    if (tid == 0) {
        if (iters_out) iters_out[p] = it;
        if (status) status[p] = st;
    }
}

template <int NS>
int lqr_block_launch(sopot_b200_ctx* ctx, bool strict, const double* A, const double* B, size_t P, const double* dQ, double R,
                     double dt, int max_iter, double tol, double* K, double* Pric, int32_t* iters, int32_t* status) {
    if (strict) lqr_block_kernel<NS, true><<<(unsigned)P, NS * NS, 0, ctx->stream>>>(A, B, P, dQ, R, dt, max_iter, tol, K, Pric, iters, status);
    else lqr_block_kernel<NS, false><<<(unsigned)P, NS * NS, 0, ctx->stream>>>(A, B, P, dQ, R, dt, max_iter, tol, K, Pric, iters, status);
    SB_CHECK_LAUNCH(ctx);
    return SOPOT_B200_OK;
}

}  // namespace

SB_EXPORT int sopot_b200_lqr4_gain(sopot_b200_ctx* ctx, const double* A, const double* B, size_t P, const double* Q16,
                                   double R, double riccati_dt, size_t max_iter, double tol,
                                   const sopot_b200_rk4_opts* opts, double* K, double* Pric, int32_t* iters,
                                   int32_t* status) {
    if (!ctx) return SOPOT_B200_EINVAL;
    if (!A || !B || !Q16 || !opts || !K) return sb_fail(ctx, SOPOT_B200_EINVAL, "NULL argument");
    if (opts->mem > 1 || opts->fp_mode > 1) return sb_fail(ctx, SOPOT_B200_EINVAL, "bad opts: mem and fp_mode must be 0 or 1");
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
        // the contract kernel compares squared norms with tol^2: below 1e-150 that underflows, use the exact kernel
        if (opts->fp_mode == SOPOT_B200_FP_STRICT || !(tol >= 1e-150))
            lqr4_kernel<true><<<grid, 128, 0, ctx->stream>>>((const double*)d_A, (const double*)d_B, P, qr, riccati_dt, (int)max_iter,
                                                             tol, (double*)d_K, (double*)d_P, (int32_t*)d_it, (int32_t*)d_st);
        else
            lqr4_kernel<false><<<grid, 128, 0, ctx->stream>>>((const double*)d_A, (const double*)d_B, P, qr, riccati_dt, (int)max_iter,
                                                              tol, (double*)d_K, (double*)d_P, (int32_t*)d_it, (int32_t*)d_st);
        SB_CHECK_LAUNCH(ctx);
    }
    return sg.finish();
}

SB_EXPORT int sopot_b200_lqr_gain(sopot_b200_ctx* ctx, int state_dim, const double* A, const double* B, size_t P, const double* Q,
                                  double R, double riccati_dt, size_t max_iter, double tol, const sopot_b200_rk4_opts* opts,
                                  double* K, double* Pric, int32_t* iters, int32_t* status) {
    if (!ctx) return SOPOT_B200_EINVAL;
    if (state_dim == 4) return sopot_b200_lqr4_gain(ctx, A, B, P, Q, R, riccati_dt, max_iter, tol, opts, K, Pric, iters, status);
    if (state_dim != 6 && state_dim != 14)
        return sb_fail(ctx, SOPOT_B200_EINVAL, "LQR kernels exist for state dimensions 4, 6 and 14 (1, 2 and 6 links)");
    if (!A || !B || !Q || !opts || !K) return sb_fail(ctx, SOPOT_B200_EINVAL, "NULL argument");
    if (opts->mem > 1 || opts->fp_mode > 1) return sb_fail(ctx, SOPOT_B200_EINVAL, "bad opts: mem and fp_mode must be 0 or 1");
    if (opts->mem > 1 || opts->fp_mode > 1) return sb_fail(ctx, SOPOT_B200_EINVAL, "bad opts");
    if (!(riccati_dt > 0.0) || max_iter > 0x7fffffff || P > 0x7fffffff) return sb_fail(ctx, SOPOT_B200_EINVAL, "bad riccati_dt / max_iter / P");
    const size_t ns = (size_t)state_dim, nn = ns * ns;
    Staging sg(ctx, opts->mem);
    int rc;
    if (sg.host && (rc = sb_arena_begin(ctx, 2 * sb_align(nn * P * 8) + 2 * sb_align(ns * P * 8) + 2 * sb_align(P * 4) + 4096)))
        return rc;
    const void *d_A, *d_B; void *d_K, *d_P, *d_it, *d_st;
    if ((rc = sg.in(A, nn * P * 8, &d_A))) return rc;
    if ((rc = sg.in(B, ns * P * 8, &d_B))) return rc;
    if ((rc = sg.out(K, ns * P * 8, &d_K))) return rc;
    if ((rc = sg.out(Pric, nn * P * 8, &d_P))) return rc;
    if ((rc = sg.out(iters, P * 4, &d_it))) return rc;
    if ((rc = sg.out(status, P * 4, &d_st))) return rc;
    void* d_Q;                                             // the weights are HOST values; they travel through the scratch
    if ((rc = sb_scratch(ctx, nn * 8 + 256, &d_Q))) return rc;
    SB_CUDA(ctx, cudaMemcpyAsync(d_Q, Q, nn * 8, cudaMemcpyHostToDevice, ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // Q may be a temporary of the caller
    if (P) {
        const bool strict = opts->fp_mode == SOPOT_B200_FP_STRICT;
        if (state_dim == 6) rc = lqr_block_launch<6>(ctx, strict, (const double*)d_A, (const double*)d_B, P, (const double*)d_Q, R, riccati_dt,
                                                     (int)max_iter, tol, (double*)d_K, (double*)d_P, (int32_t*)d_it, (int32_t*)d_st);
        else rc = lqr_block_launch<14>(ctx, strict, (const double*)d_A, (const double*)d_B, P, (const double*)d_Q, R, riccati_dt,
                                       (int)max_iter, tol, (double*)d_K, (double*)d_P, (int32_t*)d_it, (int32_t*)d_st);
        if (rc) return rc;
    }
    return sg.finish();
}
