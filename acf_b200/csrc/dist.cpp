// dist.cpp -- run-time binding of NCCL (see dist.h)
#include "dist.h"
#include <dlfcn.h>
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

} // namespace acfb
