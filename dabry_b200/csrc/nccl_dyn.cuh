// NCCL, bound at run time.  The library has no link-time dependency on libnccl: the first dabry_comm_* call looks for
// an NCCL already loaded into the process (the host application's - e.g. the one bundled with PyTorch - so that a
// process never runs two copies), then for libnccl.so.2 on the loader path.  Only the handful of entry points the
// solver's exchanges need are bound (ring send/recv inside one group, all-reduce(min), all-gather).
// Types are restated from the public nccl.h (stable since NCCL 2.4): ncclUniqueId = 128 opaque bytes passed by value.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <mutex>
#include <string>

namespace dabry {
namespace nccl {

typedef struct ncclComm* Comm_t;
struct UniqueId {
    char internal[128];
};
enum DataType : int { kInt8 = 0, kUint8 = 1, kInt32 = 2, kUint32 = 3, kInt64 = 4, kUint64 = 5, kFloat64 = 8 };
enum RedOp : int { kSum = 0, kProd = 1, kMax = 2, kMin = 3 };

struct Api {
    int (*GetUniqueId)(UniqueId*) = nullptr;
    int (*CommInitRank)(Comm_t*, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(Comm_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, Comm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, Comm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, Comm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, Comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    std::string error;  // empty: bound
    bool ok() const { return error.empty(); }
};

inline const Api& api() {
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // the process's own copy first
        if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) {
            a.error = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "");
            return;
        }
        bool all = true;
        auto bind = [&](auto& fn, const char* name) {
            fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(dlsym(lib, name));
            if (!fn) {
                all = false;
                a.error = std::string("NCCL symbol missing: ") + name;
            }
        };
        bind(a.GetUniqueId, "ncclGetUniqueId");
        bind(a.CommInitRank, "ncclCommInitRank");
        bind(a.CommDestroy, "ncclCommDestroy");
        bind(a.Send, "ncclSend");
        bind(a.Recv, "ncclRecv");
        bind(a.AllReduce, "ncclAllReduce");
        bind(a.AllGather, "ncclAllGather");
        bind(a.GroupStart, "ncclGroupStart");
        bind(a.GroupEnd, "ncclGroupEnd");
        bind(a.GetErrorString, "ncclGetErrorString");
        bind(a.GetVersion, "ncclGetVersion");
        (void) all;
    });
    return a;
}

}  // namespace nccl
}  // namespace dabry
