// context.cu -- ctx lifecycle, memory, events, staging arena and the NCCL binding of libsopot_b200.so.
#include "common.cuh"
#include "nccl_api.h"

#include <cstdarg>
#include <dlfcn.h>
#include <mutex>

static thread_local std::string g_init_error;   // per host thread: sopot_b200_init may run concurrently, one ctx per thread

int sb_fail(sopot_b200_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->error = buf; else g_init_error = buf;
    return code;
}

int sb_cuda_fail(sopot_b200_ctx* ctx, cudaError_t e, const char* what, const char* file, int line) {
    return sb_fail(ctx, SOPOT_B200_ECUDA, "CUDA error %d (%s) at %s:%d in `%s`", (int)e,
                   cudaGetErrorString(e), file, line, what);
}

// ---- NCCL through dlopen: the library must not depend on a particular libnccl at link time ----
static NcclApi g_nccl;
static std::once_flag g_nccl_once;
static std::string g_nccl_error;

static void nccl_load() {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) { g_nccl_error = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return; }
#define SB_SYM(field, name)                                                              \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));             \
    if (!g_nccl.field) { g_nccl_error = "missing NCCL symbol " name; return; }
    SB_SYM(GetUniqueId, "ncclGetUniqueId")
    SB_SYM(CommInitRank, "ncclCommInitRank")
    SB_SYM(CommDestroy, "ncclCommDestroy")
    SB_SYM(AllReduce, "ncclAllReduce")
    SB_SYM(Send, "ncclSend")
    SB_SYM(Recv, "ncclRecv")
    SB_SYM(GroupStart, "ncclGroupStart")
    SB_SYM(GroupEnd, "ncclGroupEnd")
    SB_SYM(GetErrorString, "ncclGetErrorString")
#undef SB_SYM
    g_nccl.loaded = true;
}

const NcclApi* sb_nccl(std::string* err) {
    std::call_once(g_nccl_once, nccl_load);
    if (!g_nccl.loaded) { if (err) *err = g_nccl_error; return nullptr; }
    return &g_nccl;
}

SB_EXPORT int sopot_b200_abi_version(void) { return SOPOT_B200_ABI_VERSION; }

SB_EXPORT int sopot_b200_nccl_unique_id(void* out128) {
    std::string err;
    const NcclApi* api = sb_nccl(&err);
    if (!api) return sb_fail(nullptr, SOPOT_B200_ENCCL, "%s", err.c_str());
    SbNcclUniqueId id;
    int rc = api->GetUniqueId(&id);
    if (rc != 0) return sb_fail(nullptr, SOPOT_B200_ENCCL, "ncclGetUniqueId: %s", api->GetErrorString(rc));
    memcpy(out128, &id, sizeof id);
    return SOPOT_B200_OK;
}

SB_EXPORT int sopot_b200_init(sopot_b200_ctx** out, int device, const void* nccl_unique_id, int rank,
                              int nranks) {
    if (!out) return sb_fail(nullptr, SOPOT_B200_EINVAL, "out is NULL");
    *out = nullptr;
    if (nranks < 1 || rank < 0 || rank >= nranks)
        return sb_fail(nullptr, SOPOT_B200_EINVAL, "bad rank %d / nranks %d", rank, nranks);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return sb_fail(nullptr, SOPOT_B200_ENODEVICE,
                       "no CUDA device visible (%s); libsopot_b200 has no CPU fallback",
                       e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count)
        return sb_fail(nullptr, SOPOT_B200_EINVAL, "device %d out of range (count %d)", device, count);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return sb_cuda_fail(nullptr, e, "cudaGetDeviceProperties", __FILE__, __LINE__);
    if (prop.major != 10)
        return sb_fail(nullptr, SOPOT_B200_ENODEVICE,
                       "device %d is sm_%d%d; this library carries sm_100a code only", device,
                       prop.major, prop.minor);
    if ((e = cudaSetDevice(device)) != cudaSuccess)
        return sb_cuda_fail(nullptr, e, "cudaSetDevice", __FILE__, __LINE__);

    auto* ctx = new sopot_b200_ctx();
    ctx->device = device;
    ctx->rank = rank;
    ctx->nranks = nranks;
    ctx->sm_count = prop.multiProcessorCount;
    auto bail = [&](int code) { sopot_b200_destroy(ctx); return code; };
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
        g_init_error = "cudaStreamCreate failed";
        return bail(SOPOT_B200_ECUDA);
    }
    bool ev_ok = true;
    for (auto& ev : ctx->events) ev_ok &= cudaEventCreate(&ev) == cudaSuccess;
    ev_ok &= cudaEventCreateWithFlags(&ctx->ev_a, cudaEventDisableTiming) == cudaSuccess;
    ev_ok &= cudaEventCreateWithFlags(&ctx->ev_b, cudaEventDisableTiming) == cudaSuccess;
    if (!ev_ok) {
        g_init_error = "cudaEventCreate failed";
        return bail(SOPOT_B200_ECUDA);
    }

    if (nranks > 1) {
        if (!nccl_unique_id) {
            g_init_error = "nranks > 1 needs the 128-byte id from sopot_b200_nccl_unique_id()";
            return bail(SOPOT_B200_EINVAL);
        }
        std::string err;
        const NcclApi* api = sb_nccl(&err);
        if (!api) { g_init_error = err; return bail(SOPOT_B200_ENCCL); }
        SbNcclUniqueId id;
        memcpy(&id, nccl_unique_id, sizeof id);
        int rc = api->CommInitRank(&ctx->nccl_comm, nranks, id, rank);
        if (rc != 0) {
            g_init_error = std::string("ncclCommInitRank: ") + api->GetErrorString(rc);
            ctx->nccl_comm = nullptr;
            return bail(SOPOT_B200_ENCCL);
        }
        ctx->nccl = api;
    }
    *out = ctx;
    return SOPOT_B200_OK;
}

SB_EXPORT void sopot_b200_destroy(sopot_b200_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (int k = 0; k < 2; ++k) {
        if (ctx->p2p.peer_ghost[k]) cudaIpcCloseMemHandle(ctx->p2p.peer_ghost[k]);
        if (ctx->p2p.peer_flags[k]) cudaIpcCloseMemHandle(ctx->p2p.peer_flags[k]);
    }
    if (ctx->p2p.ghost) cudaFree(ctx->p2p.ghost);
    if (ctx->p2p.flags) cudaFree(ctx->p2p.flags);
    if (ctx->nccl_comm && ctx->nccl) ctx->nccl->CommDestroy(ctx->nccl_comm);
    if (ctx->arena) cudaFree(ctx->arena);
    if (ctx->scratch) cudaFree(ctx->scratch);
    for (auto& ev : ctx->events) if (ev) cudaEventDestroy(ev);
    if (ctx->ev_a) cudaEventDestroy(ctx->ev_a);
    if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
    if (ctx->stream && ctx->own_stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
}

SB_EXPORT const char* sopot_b200_last_error(const sopot_b200_ctx* ctx) {
    return ctx ? ctx->error.c_str() : g_init_error.c_str();
}

SB_EXPORT int sopot_b200_sync(sopot_b200_ctx* ctx) {
    if (!ctx) return SOPOT_B200_EINVAL;
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->p2p.flags) {
        // a kernel of the peer-memory halo exchange gave up waiting for a neighbouring rank (grid2d.cu, p2p_wait)
        unsigned err = 0;
        SB_CUDA(ctx, cudaMemcpy(&err, ctx->p2p.flags + 96, sizeof err, cudaMemcpyDeviceToHost));
        if (err) {
            cudaMemset(ctx->p2p.flags + 96, 0, sizeof err);
            return sb_fail(ctx, SOPOT_B200_ENCCL, "halo exchange over peer memory timed out waiting for a neighbouring rank (its results are invalid)");
        }
    }
    return SOPOT_B200_OK;
}

SB_EXPORT int sopot_b200_get_stream(sopot_b200_ctx* ctx, void** cuda_stream) {
    if (!ctx || !cuda_stream) return SOPOT_B200_EINVAL;
    *cuda_stream = ctx->stream;
    return SOPOT_B200_OK;
}

SB_EXPORT int sopot_b200_note_launches(sopot_b200_ctx* ctx, uint64_t count) {
    if (!ctx) return SOPOT_B200_EINVAL;
    ctx->launches += count;
    return SOPOT_B200_OK;
}

SB_EXPORT int sopot_b200_set_stream(sopot_b200_ctx* ctx, void* cuda_stream) {
    if (!ctx) return SOPOT_B200_EINVAL;
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (cuda_stream) {
        ctx->stream = static_cast<cudaStream_t>(cuda_stream);
        ctx->own_stream = false;
    } else {
        SB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    return SOPOT_B200_OK;
}

SB_EXPORT int sopot_b200_device_info(sopot_b200_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor,
                                     size_t* global_mem_bytes, char* name, size_t name_len) {
    if (!ctx) return SOPOT_B200_EINVAL;
    cudaDeviceProp prop;
    SB_CUDA(ctx, cudaGetDeviceProperties(&prop, ctx->device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (global_mem_bytes) *global_mem_bytes = prop.totalGlobalMem;
    if (name && name_len) { strncpy(name, prop.name, name_len - 1); name[name_len - 1] = 0; }
    return SOPOT_B200_OK;
}

// ---- memory ----------------------------------------------------------------------------------
SB_EXPORT int sopot_b200_malloc(sopot_b200_ctx* ctx, size_t bytes, void** dptr) {
    if (!ctx || !dptr) return SOPOT_B200_EINVAL;
    SB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 1);
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return sb_fail(ctx, SOPOT_B200_ENOMEM, "cudaMalloc(%zu) failed", bytes); }
    SB_CUDA(ctx, e);
    return SOPOT_B200_OK;
}
SB_EXPORT int sopot_b200_free(sopot_b200_ctx* ctx, void* dptr) {
    if (!ctx) return SOPOT_B200_EINVAL;
    SB_CUDA(ctx, cudaFree(dptr));
    return SOPOT_B200_OK;
}
SB_EXPORT int sopot_b200_host_alloc(sopot_b200_ctx* ctx, size_t bytes, void** hptr) {
    if (!ctx || !hptr) return SOPOT_B200_EINVAL;
    SB_CUDA(ctx, cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return SOPOT_B200_OK;
}
SB_EXPORT int sopot_b200_host_free(sopot_b200_ctx* ctx, void* hptr) {
    if (!ctx) return SOPOT_B200_EINVAL;
    SB_CUDA(ctx, cudaFreeHost(hptr));
    return SOPOT_B200_OK;
}
SB_EXPORT int sopot_b200_memcpy_h2d(sopot_b200_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx) return SOPOT_B200_EINVAL;
    SB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return SOPOT_B200_OK;
}
SB_EXPORT int sopot_b200_memcpy_d2h(sopot_b200_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx) return SOPOT_B200_EINVAL;
    SB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return SOPOT_B200_OK;
}
SB_EXPORT int sopot_b200_memcpy_d2d(sopot_b200_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx) return SOPOT_B200_EINVAL;
    SB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return SOPOT_B200_OK;
}
SB_EXPORT int sopot_b200_memset(sopot_b200_ctx* ctx, void* dst, int byte, size_t bytes) {
    if (!ctx) return SOPOT_B200_EINVAL;
    SB_CUDA(ctx, cudaMemsetAsync(dst, byte, bytes, ctx->stream));
    return SOPOT_B200_OK;
}

// ---- events ----------------------------------------------------------------------------------
SB_EXPORT int sopot_b200_event_record(sopot_b200_ctx* ctx, int slot) {
    if (!ctx || slot < 0 || slot >= 16) return SOPOT_B200_EINVAL;
    SB_CUDA(ctx, cudaEventRecord(ctx->events[slot], ctx->stream));
    return SOPOT_B200_OK;
}
SB_EXPORT int sopot_b200_event_elapsed_ms(sopot_b200_ctx* ctx, int a, int b, float* ms) {
    if (!ctx || !ms || a < 0 || a >= 16 || b < 0 || b >= 16) return SOPOT_B200_EINVAL;
    SB_CUDA(ctx, cudaEventSynchronize(ctx->events[b]));
    SB_CUDA(ctx, cudaEventElapsedTime(ms, ctx->events[a], ctx->events[b]));
    return SOPOT_B200_OK;
}
SB_EXPORT uint64_t sopot_b200_launch_count(const sopot_b200_ctx* ctx) { return ctx ? ctx->launches : 0; }

SB_EXPORT size_t sopot_b200_rk4_num_steps(double t0, double t1, double dt) {
    // the reference's loop, core/solver.hpp:91,127 (accumulated time, not t0 + n*dt)
    if (!(dt > 0.0)) return 0;
    double t = t0;
    size_t n = 0;
    while (t < t1 - dt * 0.5) { t += dt; ++n; }
    return n;
}

// ---- arena / scratch / staging -----------------------------------------------------------------
int sb_arena_begin(sopot_b200_ctx* ctx, size_t total) {
    ctx->arena_off = 0;
    if (total <= ctx->arena_cap) return SOPOT_B200_OK;
    if (ctx->arena) { SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); SB_CUDA(ctx, cudaFree(ctx->arena)); }
    ctx->arena = nullptr; ctx->arena_cap = 0;
    size_t cap = total + total / 8 + 4096;
    cudaError_t e = cudaMalloc(&ctx->arena, cap);
    if (e != cudaSuccess) { cudaGetLastError(); return sb_fail(ctx, SOPOT_B200_ENOMEM, "staging arena cudaMalloc(%zu) failed", cap); }
    ctx->arena_cap = cap;
    return SOPOT_B200_OK;
}
void* sb_arena_take(sopot_b200_ctx* ctx, size_t bytes) {
    size_t off = ctx->arena_off;
    ctx->arena_off = off + sb_align(bytes);
    return ctx->arena_off <= ctx->arena_cap ? ctx->arena + off : nullptr;
}
int sb_scratch(sopot_b200_ctx* ctx, size_t bytes, void** out) {
    if (bytes > ctx->scratch_cap) {
        if (ctx->scratch) { SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); SB_CUDA(ctx, cudaFree(ctx->scratch)); }
        ctx->scratch = nullptr; ctx->scratch_cap = 0;
        cudaError_t e = cudaMalloc(&ctx->scratch, bytes);
        if (e != cudaSuccess) { cudaGetLastError(); return sb_fail(ctx, SOPOT_B200_ENOMEM, "scratch cudaMalloc(%zu) failed", bytes); }
        ctx->scratch_cap = bytes;
    }
    *out = ctx->scratch;
    return SOPOT_B200_OK;
}

int Staging::in(const void* user, size_t bytes, const void** dev) {
    if (!user || !bytes) { *dev = nullptr; return SOPOT_B200_OK; }
    if (!host) { *dev = user; return SOPOT_B200_OK; }
    void* d = sb_arena_take(ctx, bytes);
    if (!d) return sb_fail(ctx, SOPOT_B200_ENOMEM, "staging arena exhausted (internal sizing error)");
    SB_CUDA(ctx, cudaMemcpyAsync(d, user, bytes, cudaMemcpyHostToDevice, ctx->stream));
    *dev = d;
    return SOPOT_B200_OK;
}
int Staging::out(void* user, size_t bytes, void** dev) {
    if (!user || !bytes) { *dev = nullptr; return SOPOT_B200_OK; }
    if (!host) { *dev = user; return SOPOT_B200_OK; }
    void* d = sb_arena_take(ctx, bytes);
    if (!d) return sb_fail(ctx, SOPOT_B200_ENOMEM, "staging arena exhausted (internal sizing error)");
    outs.push_back({user, d, bytes});
    *dev = d;
    return SOPOT_B200_OK;
}
int Staging::finish() {
    for (auto& o : outs)
        SB_CUDA(ctx, cudaMemcpyAsync(o.user, o.dev, o.bytes, cudaMemcpyDeviceToHost, ctx->stream));
    outs.clear();
    return SOPOT_B200_OK;
}
