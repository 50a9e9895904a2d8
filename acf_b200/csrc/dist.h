// dist.h -- the handful of NCCL entry points the engine uses, bound at run time (dlopen "libnccl.so.2": the copy already
// loaded in the process -- e.g. the one torch brought -- or the system's), so libacf_b200.so carries no link-time dependency
// on NCCL and single-GPU hosts never touch it.  Types restate NCCL's public ABI (nccl.h): an opaque communicator pointer, a
// 128-byte unique id, int result codes, ncclUint8 = 1.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace acfb
{

struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;

struct NcclApi
{
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommInitAll)(NcclComm*, int, const int*) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    static const NcclApi& get(); // throws std::runtime_error when NCCL cannot be loaded
};
constexpr int kNcclUint8 = 1;

} // namespace acfb
