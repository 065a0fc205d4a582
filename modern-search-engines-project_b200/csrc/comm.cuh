// NCCL, loaded at run time.  The library has no link-time dependency on NCCL: libnccl.so.2 is opened on the first
// mse_comm_* call (a process that already loaded a copy — PyTorch ships one — gets that copy back from dlopen).
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

#include "common.cuh"

namespace mse {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

inline NcclApi& nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
#define MSE_NCCL_SYM(field, sym) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, sym))
        MSE_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
        MSE_NCCL_SYM(CommInitRank, "ncclCommInitRank");
        MSE_NCCL_SYM(CommDestroy, "ncclCommDestroy");
        MSE_NCCL_SYM(GroupStart, "ncclGroupStart");
        MSE_NCCL_SYM(GroupEnd, "ncclGroupEnd");
        MSE_NCCL_SYM(Send, "ncclSend");
        MSE_NCCL_SYM(Recv, "ncclRecv");
        MSE_NCCL_SYM(AllGather, "ncclAllGather");
        MSE_NCCL_SYM(AllReduce, "ncclAllReduce");
        MSE_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef MSE_NCCL_SYM
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Send && api.Recv &&
                 api.AllGather && api.AllReduce && api.GetErrorString;
    });
    return api;
}

#define MSE_NCCL_TRY(expr)                                                                                  \
    do {                                                                                                    \
        ncclResult_t _r = (expr);                                                                           \
        if (_r != ncclSuccess) {                                                                            \
            mse::set_error("%s failed: %s (%s:%d)", #expr, mse::nccl_api().GetErrorString(_r), __FILE__, __LINE__); \
            return MSE_ERR_COMM;                                                                            \
        }                                                                                                   \
    } while (0)

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    bool owned = false;
    bool active() const { return comm != nullptr || world == 1; }
};

// block w of `send` ([world] blocks of `bytes`) goes to rank w; block w of `recv` came from rank w
inline int comm_all_to_all(const Comm& c, const void* send, void* recv, size_t bytes, cudaStream_t st) {
    if (c.world == 1) {
        if (send != recv) MSE_CUDA_TRY(cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, st));
        return MSE_OK;
    }
    NcclApi& n = nccl_api();
    MSE_NCCL_TRY(n.GroupStart());
    for (int w = 0; w < c.world; ++w) {
        MSE_NCCL_TRY(n.Send(static_cast<const char*>(send) + size_t(w) * bytes, bytes, ncclChar, w, c.comm, st));
        MSE_NCCL_TRY(n.Recv(static_cast<char*>(recv) + size_t(w) * bytes, bytes, ncclChar, w, c.comm, st));
    }
    MSE_NCCL_TRY(n.GroupEnd());
    return MSE_OK;
}

inline int comm_all_gather(const Comm& c, const void* send, void* recv, size_t bytes, cudaStream_t st) {
    if (c.world == 1) {
        if (send != recv) MSE_CUDA_TRY(cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, st));
        return MSE_OK;
    }
    MSE_NCCL_TRY(nccl_api().AllGather(send, recv, bytes, ncclChar, c.comm, st));
    return MSE_OK;
}

inline int comm_all_reduce_min_u32(const Comm& c, void* buf, size_t count, cudaStream_t st) {
    if (c.world == 1) return MSE_OK;
    MSE_NCCL_TRY(nccl_api().AllReduce(buf, buf, count, ncclUint32, ncclMin, c.comm, st));
    return MSE_OK;
}

inline int comm_all_reduce_max_i32(const Comm& c, void* buf, size_t count, cudaStream_t st) {
    if (c.world == 1) return MSE_OK;
    MSE_NCCL_TRY(nccl_api().AllReduce(buf, buf, count, ncclInt32, ncclMax, c.comm, st));
    return MSE_OK;
}

}  // namespace mse
