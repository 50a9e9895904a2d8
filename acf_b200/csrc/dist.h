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
// ncclConfig_t as nccl.h 2.27 lays it out (newer libraries accept it: the struct carries its size and version)
struct NcclConfig
{
    size_t size; unsigned magic, version;
    int blocking, cgaClusterSize, minCTAs, maxCTAs; const char* netName; int splitShare, trafficClass; const char* commName;
    int collnetEnable, CTAPolicy, shrinkShare, nvlsCTAs;
};
NcclConfig ncclDefaultConfig();

struct NcclApi
{
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommInitAll)(NcclComm*, int, const int*) = nullptr;
    int (*CommInitRankConfig)(NcclComm*, int, NcclUniqueId, int, NcclConfig*) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    static const NcclApi& get(); // throws std::runtime_error when NCCL cannot be loaded
};
constexpr int kNcclUint8 = 1;

// ---- single-node exchange through POSIX shared memory ----------------------------------------------------------------------
// The gather moves tens of kilobytes per batch between processes (or threads) of ONE box.  A ring of per-rank slots in a shared
// segment named after the communicator's unique id carries it with two atomics per slot -- no kernel on the device (the engine's
// persistent kernels hold every SM: a collective's kernel has to wait for one and then keeps it while it waits for the slowest
// rank), no copy engine, no host <-> device bounce.  publish() blocks while the root is more than `ring` batches behind.
class ShmExchange
{
public:
    static ShmExchange* open(const unsigned char id[128], int rank, int world, size_t slotBytes); // rank 0 creates, the others attach
    ~ShmExchange();
    void publish(unsigned long long batch, const void* data, size_t bytes);
    const unsigned char* wait(unsigned long long batch, int rank, size_t* bytes); // root: rank's record of `batch`
    void consumed(unsigned long long batch);                                      // root: every slot of `batch` may be reused
    int world() const { return world_; }
private:
    ShmExchange() {}
    unsigned char* slot(unsigned long long batch, int rank) const;
    void* base_ = nullptr;
    size_t mapBytes_ = 0, slotBytes_ = 0;
    int rank_ = 0, world_ = 1;
    static constexpr int kRing = 8;
};

} // namespace acfb
