// stats_peaks.cu -- K4 dispersion statistics (device reduction + NCCL all-reduce) and the two roofline
// probes (FP64 DFMA issue rate, HBM copy bandwidth).
#include "common.cuh"
#include "nccl_api.h"
#include <cmath>
#include <vector>

namespace {

// per-row partial: count, sum, sumsq, nonfinite, min, max
struct Partial { double cnt, sum, sq, bad, mn, mx; };

__device__ __forceinline__ Partial combine(const Partial& a, const Partial& b) {
    return Partial{a.cnt + b.cnt, a.sum + b.sum, a.sq + b.sq, a.bad + b.bad, fmin(a.mn, b.mn), fmax(a.mx, b.mx)};
}
__device__ __forceinline__ Partial warp_reduce(Partial p) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Partial q;
        q.cnt = __shfl_down_sync(0xffffffffu, p.cnt, o); q.sum = __shfl_down_sync(0xffffffffu, p.sum, o);
        q.sq = __shfl_down_sync(0xffffffffu, p.sq, o);   q.bad = __shfl_down_sync(0xffffffffu, p.bad, o);
        q.mn = __shfl_down_sync(0xffffffffu, p.mn, o);   q.mx = __shfl_down_sync(0xffffffffu, p.mx, o);
        p = combine(p, q);
    }
    return p;
}
__device__ __forceinline__ Partial block_reduce(Partial p) {
    __shared__ Partial s[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    p = warp_reduce(p);
    if (lane == 0) s[w] = p;
    __syncthreads();
    if (w == 0) {
        const double inf = __longlong_as_double(0x7ff0000000000000LL);
        p = lane < (int)(blockDim.x >> 5) ? s[lane] : Partial{0, 0, 0, 0, inf, -inf};
        p = warp_reduce(p);
    }
    return p;
}

// grid (nblk, rows), block 256: partials[row][blk]
__global__ void __launch_bounds__(256)
dispersion_partial_kernel(const double* __restrict__ v, const size_t n, Partial* __restrict__ partials) {
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const double* row = v + (size_t)blockIdx.y * n;
    Partial p{0, 0, 0, 0, inf, -inf};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double x = __ldg(row + i);
        if (isfinite(x)) { p.cnt += 1.0; p.sum += x; p.sq += x * x; p.mn = fmin(p.mn, x); p.mx = fmax(p.mx, x); }
        else p.bad += 1.0;
    }
    p = block_reduce(p);
    if (threadIdx.x == 0) partials[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = p;
}

// grid (rows), block 256: out_sum[row*4 + {cnt,sum,sq,bad}], out_min[row], out_max[row]
__global__ void __launch_bounds__(256)
dispersion_final_kernel(const Partial* __restrict__ partials, const int nblk, double* __restrict__ out_sum,
                        double* __restrict__ out_min, double* __restrict__ out_max) {
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    Partial p{0, 0, 0, 0, inf, -inf};
    for (int b = threadIdx.x; b < nblk; b += blockDim.x) p = combine(p, partials[(size_t)blockIdx.x * nblk + b]);
    p = block_reduce(p);
    if (threadIdx.x == 0) {
        double* s = out_sum + 4 * blockIdx.x;
        s[0] = p.cnt; s[1] = p.sum; s[2] = p.sq; s[3] = p.bad;
        out_min[blockIdx.x] = p.mn; out_max[blockIdx.x] = p.mx;
    }
}

// Second pass: centred second moment  M2 = sum (x - mean)^2  per row, with the GLOBAL mean = sum / count taken from the
// all-reduced first-pass sums (identical bits on every rank).  sumsq / count - mean^2 on raw moments loses
// (mean / stddev)^2 * eps of relative accuracy (an apogee near 1e5 m with a 1 m spread keeps ~6 digits and can go
// negative); the centred sum keeps full precision and costs one more read of the rows (tens of microseconds).
// grid (nblk, rows), block 256: m2_partials[row][blk]
__global__ void __launch_bounds__(256)
dispersion_m2_partial_kernel(const double* __restrict__ v, const size_t n, const double* __restrict__ sums,
                             double* __restrict__ m2_partials) {
    const double* row = v + (size_t)blockIdx.y * n;
    const double cnt = sums[4 * blockIdx.y], mean = cnt > 0 ? sums[4 * blockIdx.y + 1] / cnt : 0.0;
    double acc = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double x = __ldg(row + i);
        if (isfinite(x)) { const double d = x - mean; acc = fma(d, d, acc); }
    }
    __shared__ double sh[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sh[w];
        m2_partials[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
}
// grid (rows), one warp: fixed summation order, so the result does not depend on scheduling
__global__ void dispersion_m2_final_kernel(const double* __restrict__ m2_partials, const int nblk, double* __restrict__ m2) {
    double acc = 0.0;
    for (int b = threadIdx.x; b < nblk; b += 32) acc += m2_partials[(size_t)blockIdx.x * nblk + b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) m2[blockIdx.x] = acc;
}

// ---- roofline probes ----------------------------------------------------------------------------------
// 8 independent DFMA chains per thread: pure FP64-pipe issue rate, FMA = 2 flop.  The multiplicand is a per-thread
// vector register: x = fma(x, y, U) issues every 2.02 cycles per sub-partition, the best of the operand placements
// measured in tools/microbench/fp64_operands.cu (two uniform operands: 2.22; three distinct vector registers: 3.02).
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, const int iters, const double a, const double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    const double y = a + 1e-12 * threadIdx.x;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, y, b); x1 = fma(x1, y, b); x2 = fma(x2, y, b); x3 = fma(x3, y, b);
            x4 = fma(x4, y, b); x5 = fma(x5, y, b); x6 = fma(x6, y, b); x7 = fma(x7, y, b);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void __launch_bounds__(256) copy_peak_kernel(const double2* __restrict__ src, double2* __restrict__ dst, const size_t n2) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

}  // namespace

SB_EXPORT int sopot_b200_dispersion_stats(sopot_b200_ctx* ctx, const double* values, size_t rows, size_t n_local,
                                          uint32_t mem, sopot_b200_dispersion* out) {
    if (!ctx) return SOPOT_B200_EINVAL;
    if (!out || rows == 0 || mem > 1 || (n_local && !values)) return sb_fail(ctx, SOPOT_B200_EINVAL, "bad argument");
    Staging sg(ctx, mem);
    int rc;
    if (sg.host && (rc = sb_arena_begin(ctx, sb_align(rows * n_local * 8) + 4096))) return rc;
    const void* d_v = nullptr;
    if ((rc = sg.in(values, rows * n_local * 8, &d_v))) return rc;
    const int nblk = (int)std::min<size_t>(std::max<size_t>((n_local + 255) / 256, 1), (size_t)ctx->sm_count * 4);
    void* scr;
    const size_t part_bytes = sb_align(sizeof(Partial) * rows * nblk);
    if ((rc = sb_scratch(ctx, part_bytes + sb_align(rows * 7 * 8) + 256, &scr))) return rc;
    Partial* partials = static_cast<Partial*>(scr);
    double* d_sum = reinterpret_cast<double*>(static_cast<char*>(scr) + part_bytes);
    double* d_min = d_sum + 4 * rows;
    double* d_max = d_min + rows;
    double* d_m2 = d_max + rows;
    dispersion_partial_kernel<<<dim3(nblk, (unsigned)rows), 256, 0, ctx->stream>>>((const double*)d_v, n_local, partials);
    SB_CHECK_LAUNCH(ctx);
    dispersion_final_kernel<<<(unsigned)rows, 256, 0, ctx->stream>>>(partials, nblk, d_sum, d_min, d_max);
    SB_CHECK_LAUNCH(ctx);
    if (ctx->nranks > 1) {
        // the path's only exchange: <= 7*rows doubles in four small all-reduces, latency bound
        const NcclApi* n = ctx->nccl;
        SB_NCCL(ctx, n->AllReduce(d_sum, d_sum, 4 * rows, SB_NCCL_FLOAT64, SB_NCCL_SUM, ctx->nccl_comm, ctx->stream));
        SB_NCCL(ctx, n->AllReduce(d_min, d_min, rows, SB_NCCL_FLOAT64, SB_NCCL_MIN, ctx->nccl_comm, ctx->stream));
        SB_NCCL(ctx, n->AllReduce(d_max, d_max, rows, SB_NCCL_FLOAT64, SB_NCCL_MAX, ctx->nccl_comm, ctx->stream));
    }
    // centred second moment about the global mean (the partials buffer is free again: reuse it as doubles)
    double* m2_partials = reinterpret_cast<double*>(partials);
    dispersion_m2_partial_kernel<<<dim3(nblk, (unsigned)rows), 256, 0, ctx->stream>>>((const double*)d_v, n_local, d_sum, m2_partials);
    SB_CHECK_LAUNCH(ctx);
    dispersion_m2_final_kernel<<<(unsigned)rows, 32, 0, ctx->stream>>>(m2_partials, nblk, d_m2);
    SB_CHECK_LAUNCH(ctx);
    if (ctx->nranks > 1)
        SB_NCCL(ctx, ctx->nccl->AllReduce(d_m2, d_m2, rows, SB_NCCL_FLOAT64, SB_NCCL_SUM, ctx->nccl_comm, ctx->stream));
    std::vector<double> h(7 * rows);
    SB_CUDA(ctx, cudaMemcpyAsync(h.data(), d_sum, 7 * rows * 8, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (size_t r = 0; r < rows; ++r) {
        sopot_b200_dispersion& o = out[r];
        o.count = h[4 * r]; o.sum = h[4 * r + 1]; o.sumsq = h[4 * r + 2]; o.nonfinite = h[4 * r + 3];
        o.min = h[4 * rows + r]; o.max = h[5 * rows + r];
        o.mean = o.count > 0 ? o.sum / o.count : NAN;
        o.stddev = o.count > 0 ? std::sqrt(h[6 * rows + r] / o.count) : NAN;   // population standard deviation, from M2
    }
    return SOPOT_B200_OK;
}

SB_EXPORT int sopot_b200_measure_fp64_peak(sopot_b200_ctx* ctx, double* tflops) {
    if (!ctx || !tflops) return SOPOT_B200_EINVAL;
    // short (~1 ms) bursts, best of 10: the burst figure for a kernel timed alone (B200_PROFILING.md)
    const int blocks = ctx->sm_count * 8, threads = 256, iters = 1024;
    void* scr;
    int rc = sb_scratch(ctx, (size_t)blocks * threads * 8, &scr);
    if (rc) return rc;
    cudaEvent_t e0, e1;
    SB_CUDA(ctx, cudaEventCreate(&e0));
    SB_CUDA(ctx, cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 11; ++rep) {
        SB_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        dfma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>((double*)scr, iters, 1.0000001, 1e-9);
        SB_CHECK_LAUNCH(ctx);
        SB_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        SB_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0;
        SB_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        const double flop = 2.0 * 64.0 * iters * (double)blocks * threads;
        if (rep > 0) best = std::max(best, flop / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *tflops = best;
    return SOPOT_B200_OK;
}

SB_EXPORT int sopot_b200_measure_hbm_peak(sopot_b200_ctx* ctx, size_t bytes, double* gbs) {
    if (!ctx || !gbs || bytes < (1u << 20)) return SOPOT_B200_EINVAL;
    bytes &= ~size_t(255);
    void *a = nullptr, *b = nullptr;
    if (cudaMalloc(&a, bytes) != cudaSuccess || cudaMalloc(&b, bytes) != cudaSuccess) {
        cudaGetLastError();
        if (a) cudaFree(a);
        return sb_fail(ctx, SOPOT_B200_ENOMEM, "cannot allocate 2 x %zu bytes for the copy probe", bytes);
    }
    cudaMemsetAsync(a, 1, bytes, ctx->stream);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 8; ++rep) {
        cudaEventRecord(e0, ctx->stream);
        copy_peak_kernel<<<ctx->sm_count * 16, 256, 0, ctx->stream>>>((const double2*)a, (double2*)b, bytes / 16);
        ctx->launches++;
        cudaEventRecord(e1, ctx->stream);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 1) best = std::max(best, 2.0 * (double)bytes / (ms * 1e-3) / 1e9);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(a); cudaFree(b);
    SB_CUDA(ctx, cudaGetLastError());
    *gbs = best;
    return SOPOT_B200_OK;
}
