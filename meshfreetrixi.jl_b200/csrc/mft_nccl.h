// mft_nccl.h -- run-time binding to NCCL (dlopen), so that single-GPU users of libmft_b200.so do not need
// libnccl at all.  When the host program already has NCCL loaded (e.g. torch.distributed), dlopen by soname
// returns that same copy.  Override the library path with MFT_NCCL_LIB.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>
#include <stdlib.h>
#include <string>

struct NcclUniqueId {
    char internal[128];
};

constexpr int NCCL_DOUBLE = 8;  // ncclFloat64
constexpr int NCCL_SUM = 0;

struct NcclApi {
    int (*p_getUniqueId)(NcclUniqueId *) = nullptr;
    int (*p_commInitRank)(void **, int, NcclUniqueId, int) = nullptr;
    int (*p_commDestroy)(void *) = nullptr;
    int (*p_send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*p_recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*p_groupStart)() = nullptr;
    int (*p_groupEnd)() = nullptr;
    int (*p_allGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*p_allReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    const char *(*p_getLastError)(void *) = nullptr;

    int getUniqueId(void *id128) { return p_getUniqueId(reinterpret_cast<NcclUniqueId *>(id128)); }
    int commInitRank(void **comm, int n, const void *id128, int rank)
    {
        return p_commInitRank(comm, n, *reinterpret_cast<const NcclUniqueId *>(id128), rank);
    }
    int commDestroy(void *comm) { return p_commDestroy(comm); }
    int send(const void *b, size_t cnt, int dt, int peer, void *comm, cudaStream_t s) { return p_send(b, cnt, dt, peer, comm, s); }
    int recv(void *b, size_t cnt, int dt, int peer, void *comm, cudaStream_t s) { return p_recv(b, cnt, dt, peer, comm, s); }
    int groupStart() { return p_groupStart(); }
    int groupEnd() { return p_groupEnd(); }
    int allGather(const void *s, void *r, size_t cnt, int dt, void *comm, cudaStream_t st) { return p_allGather(s, r, cnt, dt, comm, st); }
    int allReduce(const void *s, void *r, size_t cnt, int dt, int op, void *comm, cudaStream_t st)
    {
        return p_allReduce(s, r, cnt, dt, op, comm, st);
    }
    const char *lastError(void *comm) { return p_getLastError ? p_getLastError(comm) : "?"; }
};

inline std::string &nccl_err_storage()
{
    static std::string e;
    return e;
}
inline const char *nccl_load_error() { return nccl_err_storage().c_str(); }

inline NcclApi *nccl_api()
{
    static NcclApi api;
    static int state = 0;  // 0 untried, 1 ok, -1 failed
    if (state == 1) return &api;
    if (state == -1) return nullptr;
    const char *env = getenv("MFT_NCCL_LIB");
    const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *nm : names) {
        if (!nm) continue;
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        nccl_err_storage() = dlerror() ? dlerror() : "dlopen(libnccl.so.2) failed";
        state = -1;
        return nullptr;
    }
    bool ok = true;
    auto sym = [&](const char *n) -> void * {
        void *p = dlsym(h, n);
        if (!p) {
            ok = false;
            nccl_err_storage() = std::string("missing symbol ") + n;
        }
        return p;
    };
    api.p_getUniqueId = reinterpret_cast<decltype(api.p_getUniqueId)>(sym("ncclGetUniqueId"));
    api.p_commInitRank = reinterpret_cast<decltype(api.p_commInitRank)>(sym("ncclCommInitRank"));
    api.p_commDestroy = reinterpret_cast<decltype(api.p_commDestroy)>(sym("ncclCommDestroy"));
    api.p_send = reinterpret_cast<decltype(api.p_send)>(sym("ncclSend"));
    api.p_recv = reinterpret_cast<decltype(api.p_recv)>(sym("ncclRecv"));
    api.p_groupStart = reinterpret_cast<decltype(api.p_groupStart)>(sym("ncclGroupStart"));
    api.p_groupEnd = reinterpret_cast<decltype(api.p_groupEnd)>(sym("ncclGroupEnd"));
    api.p_allGather = reinterpret_cast<decltype(api.p_allGather)>(sym("ncclAllGather"));
    api.p_allReduce = reinterpret_cast<decltype(api.p_allReduce)>(sym("ncclAllReduce"));
    api.p_getLastError = reinterpret_cast<decltype(api.p_getLastError)>(dlsym(h, "ncclGetLastError"));
    state = ok ? 1 : -1;
    return ok ? &api : nullptr;
}
