// nccl_api.h -- the handful of NCCL entry points libsopot_b200.so uses, bound with dlopen()/dlsym()
// at run time (context.cu) so the library loads on a box without NCCL and picks up whichever
// libnccl.so.2 the host process already has (torch's bundled one under torch.distributed).
// Declarations follow the public nccl.h (2.x): ncclUniqueId is 128 opaque bytes, ncclComm_t an
// opaque pointer, ncclFloat64 = 8, ncclSum = 0, ncclMax = 2, ncclMin = 3.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <string>

struct SbNcclUniqueId { char internal[128]; };
enum { SB_NCCL_FLOAT64 = 8, SB_NCCL_SUM = 0, SB_NCCL_MAX = 2, SB_NCCL_MIN = 3 };

struct NcclApi {
    bool loaded = false;
    int (*GetUniqueId)(SbNcclUniqueId*) = nullptr;
    int (*CommInitRank)(void** comm, int nranks, SbNcclUniqueId id, int rank) = nullptr;
    int (*CommDestroy)(void* comm) = nullptr;
    int (*AllReduce)(const void* send, void* recv, size_t count, int dtype, int op, void* comm,
                     cudaStream_t stream) = nullptr;
    int (*Send)(const void* buf, size_t count, int dtype, int peer, void* comm, cudaStream_t stream) = nullptr;
    int (*Recv)(void* buf, size_t count, int dtype, int peer, void* comm, cudaStream_t stream) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

const NcclApi* sb_nccl(std::string* err);

#define SB_NCCL(ctx, expr)                                                                     \
    do {                                                                                       \
        int _rc = (expr);                                                                      \
        if (_rc != 0)                                                                          \
            return sb_fail((ctx), SOPOT_B200_ENCCL, "NCCL error %d (%s) in `%s`", _rc,         \
                           (ctx)->nccl->GetErrorString(_rc), #expr);                           \
    } while (0)
