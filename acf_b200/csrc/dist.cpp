// dist.cpp -- run-time binding of NCCL (see dist.h)
#include "dist.h"
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <climits>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>

namespace acfb
{

const NcclApi& NcclApi::get()
{
    static NcclApi api;
    static std::once_flag once;
    static std::string err;
    std::call_once(once, [] {
        void* h = nullptr;
        for (const char* name : { "libnccl.so.2", "libnccl.so" })
        {
            h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) { err = std::string("NCCL: cannot load libnccl.so.2 (") + dlerror() + ")"; return; }
        auto bind = [&](const char* sym) -> void* {
            void* p = dlsym(h, sym);
            if (!p && err.empty()) err = std::string("NCCL: symbol ") + sym + " is missing";
            return p;
        };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(bind("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(bind("ncclCommInitRank"));
        api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(bind("ncclCommInitAll"));
        api.CommInitRankConfig = reinterpret_cast<decltype(api.CommInitRankConfig)>(bind("ncclCommInitRankConfig"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(bind("ncclCommDestroy"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(bind("ncclAllGather"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(bind("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(bind("ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(bind("ncclGetErrorString"));
        api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(bind("ncclGetVersion"));
    });
    if (!err.empty()) throw std::runtime_error(err);
    return api;
}

NcclConfig ncclDefaultConfig()
{
    NcclConfig c;
    c.size = sizeof(NcclConfig); c.magic = 0xcafebeefu; c.version = 22703u;
    c.blocking = c.cgaClusterSize = c.minCTAs = c.maxCTAs = c.splitShare = c.trafficClass = INT_MIN;
    c.collnetEnable = c.CTAPolicy = c.shrinkShare = c.nvlsCTAs = INT_MIN;
    c.netName = nullptr; c.commName = nullptr;
    return c;
}

// ---- ShmExchange -----------------------------------------------------------------------------------------------------------
namespace
{
struct ShmHeader
{
    std::atomic<unsigned> ready;          // set by the creator once the header is initialised
    std::atomic<unsigned> attached;       // ranks that have mapped the segment; the creator unlinks the name when all have
    unsigned world, ring;
    unsigned long long slotBytes;
    std::atomic<unsigned long long> consumed; // batches the root has taken
    // followed by: std::atomic<unsigned long long> seq[ring * world]  (batch + 1 once the slot holds that batch),
    //              unsigned long long bytes[ring * world], then the slots
};
std::atomic<unsigned long long>* seqOf(void* base) { return reinterpret_cast<std::atomic<unsigned long long>*>(static_cast<char*>(base) + 256); }
template <class F>
void spinUntil(F done, const char* what)
{
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned it = 0; !done(); it++)
    {
        if (it > 2000) std::this_thread::sleep_for(std::chrono::microseconds(20)); else std::this_thread::yield();
        if ((it & 1023) == 1023 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120))
            throw std::runtime_error(std::string("engine: timed out waiting for ") + what + " (a rank stopped, or the ranks run different batch sequences)");
    }
}
} // namespace

ShmExchange* ShmExchange::open(const unsigned char id[128], int rank, int world, size_t slotBytes)
{
    char name[64];
    unsigned long long h0 = 1469598103934665603ull, h1 = 0x9e3779b97f4a7c15ull;
    for (int i = 0; i < 128; i++) { h0 = (h0 ^ id[i]) * 1099511628211ull; h1 = (h1 + id[i]) * 0xff51afd7ed558ccdull; h1 ^= h1 >> 29; }
    snprintf(name, sizeof(name), "/acfb_%016llx%016llx", h0, h1);
    std::unique_ptr<ShmExchange> x(new ShmExchange());
    x->rank_ = rank; x->world_ = world;
    x->slotBytes_ = (slotBytes + 63) & ~(size_t)63;
    const size_t tables = (size_t)kRing * world * 16;
    x->mapBytes_ = 256 + ((tables + 255) & ~(size_t)255) + (size_t)kRing * world * x->slotBytes_;
    int fd = -1;
    if (rank == 0)
    {
        shm_unlink(name);
        fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, (off_t)x->mapBytes_) != 0) { if (fd >= 0) close(fd); throw std::runtime_error("engine: cannot create the shared-memory exchange segment"); }
    }
    else
    {
        const auto t0 = std::chrono::steady_clock::now();
        for (;;)
        {
            fd = shm_open(name, O_RDWR, 0600);
            struct stat st;
            if (fd >= 0 && fstat(fd, &st) == 0 && (size_t)st.st_size >= x->mapBytes_) break;
            if (fd >= 0) { close(fd); fd = -1; }
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(60)) throw std::runtime_error("engine: rank 0 never created the shared-memory exchange segment");
            std::this_thread::sleep_for(std::chrono::milliseconds(1));
        }
    }
    x->base_ = mmap(nullptr, x->mapBytes_, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (x->base_ == MAP_FAILED) { x->base_ = nullptr; throw std::runtime_error("engine: cannot map the shared-memory exchange segment"); }
    ShmHeader* H = static_cast<ShmHeader*>(x->base_);
    if (rank == 0)
    {   // a fresh segment is zero-filled: every atomic starts at 0
        H->world = (unsigned)world; H->ring = kRing; H->slotBytes = x->slotBytes_;
        H->ready.store(1, std::memory_order_release);
    }
    else
    {
        spinUntil([&] { return H->ready.load(std::memory_order_acquire) == 1; }, "the exchange segment's header");
        if (H->world != (unsigned)world || H->slotBytes != x->slotBytes_) throw std::runtime_error("engine: the ranks disagree on world size or batch capacity");
    }
    // the last rank to attach removes the name; the mapping lives on until the last rank unmaps it
    if (H->attached.fetch_add(1, std::memory_order_acq_rel) + 1 == (unsigned)world) shm_unlink(name);
    return x.release();
}

ShmExchange::~ShmExchange() { if (base_) munmap(base_, mapBytes_); }

unsigned char* ShmExchange::slot(unsigned long long batch, int rank) const
{
    const size_t tables = ((size_t)kRing * world_ * 16 + 255) & ~(size_t)255;
    return static_cast<unsigned char*>(base_) + 256 + tables + ((size_t)(batch % kRing) * world_ + rank) * slotBytes_;
}

void ShmExchange::publish(unsigned long long batch, const void* data, size_t bytes)
{
    if (bytes > slotBytes_) throw std::runtime_error("engine: detection record larger than an exchange slot");
    ShmHeader* H = static_cast<ShmHeader*>(base_);
    if (batch >= (unsigned long long)kRing)
        spinUntil([&] { return H->consumed.load(std::memory_order_acquire) + kRing > batch; }, "rank 0 to collect earlier batches");
    const size_t i = (size_t)(batch % kRing) * world_ + rank_;
    memcpy(slot(batch, rank_), data, bytes);
    reinterpret_cast<unsigned long long*>(seqOf(base_) + (size_t)kRing * world_)[i] = bytes;
    seqOf(base_)[i].store(batch + 1, std::memory_order_release);
}

const unsigned char* ShmExchange::wait(unsigned long long batch, int rank, size_t* bytes)
{
    const size_t i = (size_t)(batch % kRing) * world_ + rank;
    spinUntil([&] { return seqOf(base_)[i].load(std::memory_order_acquire) == batch + 1; }, "a rank's boxes");
    if (bytes) *bytes = reinterpret_cast<unsigned long long*>(seqOf(base_) + (size_t)kRing * world_)[i];
    return slot(batch, rank);
}

void ShmExchange::consumed(unsigned long long batch) { static_cast<ShmHeader*>(base_)->consumed.store(batch + 1, std::memory_order_release); }

} // namespace acfb
