// common.cuh -- context, error plumbing, host<->device staging and the Real<Strict> scalar used by
// every kernel of libsopot_b200.so.  sm_100a only; there is no CPU path in this library.
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/sopot_b200.h"
#include "ieee.cuh"

#define SB_EXPORT extern "C" __attribute__((visibility("default")))

struct NcclApi;   // context.cu

struct sopot_b200_ctx {
    int device = 0;
    int rank = 0, nranks = 1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    cudaStream_t copy_stream = nullptr;          // halo / overlap work
    cudaEvent_t events[16] = {};
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;  // internal cross-stream ordering
    uint64_t launches = 0;
    uint32_t ugrid_kernel = 0;                   // SOPOT_B200_FLAG_UGRID_MARCH / _GATHER of the running ugrid call (0: by size)
    std::string error;
    // grow-only device arena used to stage HOST-pointer calls
    char* arena = nullptr;
    size_t arena_cap = 0, arena_off = 0;
    // persistent scratch for device-pointer calls (grid work vectors, reduction partials)
    char* scratch = nullptr;
    size_t scratch_cap = 0;
    // NCCL (nranks > 1)
    void* nccl_comm = nullptr;
    const NcclApi* nccl = nullptr;
    // grid2d halo exchange over NVLink peer memory (grid2d.cu): ghost rows and sequence flags that the NEIGHBOURING ranks'
    // kernels write directly, mapped into their address spaces with CUDA IPC
    struct P2P {
        bool tried = false, ready = false;
        size_t slot_doubles = 0;                       // capacity of one ghost slot (4 boundary rows)
        double* ghost = nullptr;                       // [2 parity][2 side: from-up, from-down][slot_doubles]
        unsigned* flags = nullptr;                     // [0] from the rank above, [32] from the rank below, [64..65] CTA counters, [96] error
        double* peer_ghost[2] = {nullptr, nullptr};    // the rank above / below: its `ghost`
        unsigned* peer_flags[2] = {nullptr, nullptr};  // ... and its `flags`
        unsigned seq = 0;                              // exchanges completed (same on every rank)
    } p2p;
};

int sb_fail(sopot_b200_ctx* ctx, int code, const char* fmt, ...);
int sb_cuda_fail(sopot_b200_ctx* ctx, cudaError_t e, const char* what, const char* file, int line);

#define SB_CUDA(ctx, expr)                                                                  \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) return sb_cuda_fail((ctx), _e, #expr, __FILE__, __LINE__);   \
    } while (0)

#define SB_CHECK_LAUNCH(ctx)                                                                \
    do {                                                                                    \
        (ctx)->launches++;                                                                  \
        cudaError_t _e = cudaGetLastError();                                                \
        if (_e != cudaSuccess) return sb_cuda_fail((ctx), _e, "kernel launch", __FILE__, __LINE__); \
    } while (0)

// ---- arena staging -------------------------------------------------------------------------
int sb_arena_begin(sopot_b200_ctx* ctx, size_t total_bytes);          // ensure capacity, reset bump
void* sb_arena_take(sopot_b200_ctx* ctx, size_t bytes);               // 256 B aligned bump allocation
int sb_scratch(sopot_b200_ctx* ctx, size_t bytes, void** out);        // persistent scratch >= bytes
inline size_t sb_align(size_t b) { return (b + 255) & ~size_t(255); }

// A call's view of its buffers: for MEM_DEVICE the user pointers are used as they are; for MEM_HOST
// inputs are copied into the arena up front and outputs are copied back by finish().
struct Staging {
    sopot_b200_ctx* ctx;
    bool host;
    struct Out { void* user; void* dev; size_t bytes; };
    std::vector<Out> outs;
    // every entry point builds a Staging first: make the ctx's device current (a process may hold contexts on several GPUs)
    Staging(sopot_b200_ctx* c, uint32_t mem) : ctx(c), host(mem == SOPOT_B200_MEM_HOST) { if (c) cudaSetDevice(c->device); }
    // returns the device address to read from (nullptr stays nullptr)
    int in(const void* user, size_t bytes, const void** dev);
    int out(void* user, size_t bytes, void** dev);
    int finish();
};

// ---- per-instance parameter access -----------------------------------------------------------
// A parameter is an array of n doubles in the call's memory space, or (broadcast) ONE double that the entry point
// read from HOST memory and passes by value: it then lives in a uniform register for the whole kernel.
struct ParamRef {
    const double* p;   // nullptr: broadcast
    double value;
    __device__ __forceinline__ double at(size_t i) const { return p ? __ldg(p + i) : value; }
};

// ---- Real<Strict> ----------------------------------------------------------------------------
// Strict = true : every + - * is a separately rounded IEEE operation (__dadd_rn/__dmul_rn are never
//                 contracted into DFMA), divisions and sqrt are the IEEE-correct CUDA defaults: the
//                 arithmetic of the reference's g++ -O3 build without -march (no FMA).
// Strict = false: plain operators; nvcc (-fmad=true) contracts a*b+c into DFMA.
template <bool Strict>
struct Real {
    double v;
    __host__ __device__ __forceinline__ Real() : v(0.0) {}
    __host__ __device__ __forceinline__ Real(double x) : v(x) {}
    __host__ __device__ __forceinline__ explicit operator double() const { return v; }
};

// SB_COUNT_OPS (tools/rocket_opcount.py builds one translation unit with it): every strict-mode operation of thread 0 of
// block 0 is tallied by kind -- the "counting scalar" of SURVEY.md section 8d applied to the engine's OWN derivative, on the
// device, in the arithmetic the reference uses (one count per + - * / sqrt and per transcendental call).
#if defined(SB_COUNT_OPS)
static __device__ unsigned long long g_sb_ops[8];     // 0 add/sub, 1 mul, 2 div, 3 sqrt, 4 transcendental
#endif
#if defined(SB_COUNT_OPS) && defined(__CUDA_ARCH__)
__device__ __forceinline__ void sb_count(int k) { if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(&g_sb_ops[k], 1ull); }
#define SB_ADD(a, b) (sb_count(0), __dadd_rn((a), (b)))
#define SB_MUL(a, b) (sb_count(1), __dmul_rn((a), (b)))
#define SB_DIV(a, b) (sb_count(2), sb_div_rn((a), (b)))
#define SB_SQRT(a) (sb_count(3), sb_sqrt_rn((a)))
#define SB_COUNT_TRANSCENDENTAL() sb_count(4)
#define SB_COUNT_DIV() sb_count(2)
#elif defined(__CUDA_ARCH__)
#define SB_COUNT_TRANSCENDENTAL() ((void)0)
#define SB_COUNT_DIV() ((void)0)
#define SB_ADD(a, b) __dadd_rn((a), (b))
#define SB_MUL(a, b) __dmul_rn((a), (b))
#define SB_DIV(a, b) sb_div_rn((a), (b))
#define SB_SQRT(a) sb_sqrt_rn((a))
#else
#define SB_COUNT_TRANSCENDENTAL() ((void)0)
#define SB_COUNT_DIV() ((void)0)
#define SB_ADD(a, b) ((a) + (b))
#define SB_MUL(a, b) ((a) * (b))
#define SB_DIV(a, b) ((a) / (b))
#define SB_SQRT(a) sqrt((a))
#endif

template <bool S> __host__ __device__ __forceinline__ Real<S> operator+(Real<S> a, Real<S> b) {
    if constexpr (S) return Real<S>(SB_ADD(a.v, b.v)); else return Real<S>(a.v + b.v);
}
template <bool S> __host__ __device__ __forceinline__ Real<S> operator-(Real<S> a, Real<S> b) {
    if constexpr (S) return Real<S>(SB_ADD(a.v, -b.v)); else return Real<S>(a.v - b.v);
}
template <bool S> __host__ __device__ __forceinline__ Real<S> operator*(Real<S> a, Real<S> b) {
    if constexpr (S) return Real<S>(SB_MUL(a.v, b.v)); else return Real<S>(a.v * b.v);
}
template <bool S> __host__ __device__ __forceinline__ Real<S> operator/(Real<S> a, Real<S> b) {
    return Real<S>(SB_DIV(a.v, b.v));   // correctly rounded division in both modes (ieee.cuh)
}
template <bool S> __host__ __device__ __forceinline__ Real<S> operator-(Real<S> a) { return Real<S>(-a.v); }
template <bool S> __host__ __device__ __forceinline__ Real<S>& operator+=(Real<S>& a, Real<S> b) { a = a + b; return a; }
template <bool S> __host__ __device__ __forceinline__ Real<S>& operator-=(Real<S>& a, Real<S> b) { a = a - b; return a; }
template <bool S> __host__ __device__ __forceinline__ Real<S>& operator*=(Real<S>& a, Real<S> b) { a = a * b; return a; }
// mixed with double literals
#define SB_MIXED(op)                                                                               \
    template <bool S> __host__ __device__ __forceinline__ Real<S> operator op(Real<S> a, double b) { return a op Real<S>(b); } \
    template <bool S> __host__ __device__ __forceinline__ Real<S> operator op(double a, Real<S> b) { return Real<S>(a) op b; }
SB_MIXED(+) SB_MIXED(-) SB_MIXED(*) SB_MIXED(/)
#undef SB_MIXED
template <bool S> __host__ __device__ __forceinline__ bool operator<(Real<S> a, Real<S> b) { return a.v < b.v; }
template <bool S> __host__ __device__ __forceinline__ bool operator>(Real<S> a, Real<S> b) { return a.v > b.v; }
template <bool S> __host__ __device__ __forceinline__ bool operator<=(Real<S> a, Real<S> b) { return a.v <= b.v; }
template <bool S> __host__ __device__ __forceinline__ bool operator>=(Real<S> a, Real<S> b) { return a.v >= b.v; }

template <bool S> __device__ __forceinline__ Real<S> r_sqrt(Real<S> a) { return Real<S>(SB_SQRT(a.v)); }
template <bool S> __device__ __forceinline__ Real<S> r_sin(Real<S> a) { SB_COUNT_TRANSCENDENTAL(); return Real<S>(sin(a.v)); }
template <bool S> __device__ __forceinline__ Real<S> r_cos(Real<S> a) { SB_COUNT_TRANSCENDENTAL(); return Real<S>(cos(a.v)); }
template <bool S> __device__ __forceinline__ void r_sincos(Real<S> a, Real<S>& s, Real<S>& c) {
    SB_COUNT_TRANSCENDENTAL(); SB_COUNT_TRANSCENDENTAL();
    double sv, cv; sincos(a.v, &sv, &cv); s = Real<S>(sv); c = Real<S>(cv);
}
// x / d where the reference divides: IEEE division when Strict, reciprocal multiply otherwise
template <bool S> __device__ __forceinline__ Real<S> r_div_const(Real<S> a, double d) {
#ifdef __CUDA_ARCH__
    if constexpr (S) { SB_COUNT_DIV(); return Real<S>(sb_div_by(a.v, d, 1.0 / d)); } else return Real<S>(a.v * (1.0 / d));   // d is a literal: 1/d folds
#else
    if constexpr (S) return Real<S>(a.v / d); else return Real<S>(a.v * (1.0 / d));
#endif
}
// a / d for a divisor that is constant over the launch, inv = RN(1 / d) from the host
template <bool S> __device__ __forceinline__ Real<S> r_div_by(Real<S> a, double d, double inv) {
    if constexpr (S) { SB_COUNT_DIV(); return Real<S>(sb_div_by(a.v, d, inv)); } else return Real<S>(a.v * inv);
}

// The slab contract of the multi-rank grid entry points: rank r of R owns the contiguous block of rows that
// sopot_b200/distributed.py::partition assigns (sizes differ by at most one row).  Every rank derives its schedule
// (fused / overlapped / per-stage, one exchange per step or per stage) from rows / nranks alone, so the branches -- and
// with them the ncclSend/ncclRecv pairs -- agree across ranks only for exactly this decomposition: anything else is
// rejected up front instead of hanging in the first exchange.
inline void sb_partition(size_t n, int rank, int nranks, size_t* begin, size_t* count) {
    const size_t base = n / (size_t)nranks, rem = n % (size_t)nranks, r = (size_t)rank;
    *begin = r * base + (r < rem ? r : rem);
    *count = base + (r < rem ? 1 : 0);
}
inline int sb_check_slab(sopot_b200_ctx* ctx, size_t rows, size_t row_begin, size_t rows_local) {
    if (ctx->nranks <= 1) return SOPOT_B200_OK;
    size_t b, c;
    sb_partition(rows, ctx->rank, ctx->nranks, &b, &c);
    if (rows < (size_t)ctx->nranks)
        return sb_fail(ctx, SOPOT_B200_EINVAL, "%zu rows cannot be split over %d ranks", rows, ctx->nranks);
    if (row_begin != b || rows_local != c)
        return sb_fail(ctx, SOPOT_B200_EINVAL, "rank %d of %d must own rows [%zu, %zu) of %zu (the even contiguous partition), got [%zu, %zu)",
                       ctx->rank, ctx->nranks, b, b + c, rows, row_begin, row_begin + rows_local);
    return SOPOT_B200_OK;
}

inline unsigned sb_grid_for(size_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }
