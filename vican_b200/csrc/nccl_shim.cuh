// NCCL all-reduce hook for edge-sharded multi-GPU runs (one process per GPU).
//
// The library is resolved at run time (dlopen of the libnccl.so.2 that PyTorch already loaded,
// else the system one) so the extension has no link-time dependency and single-GPU use never
// touches NCCL.  Only the camera-side accumulator (n_c x 9 doubles) crosses NVLink, once per
// camera pass.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

namespace vb {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    bool ok = false;
    NcclApi() {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        GetUniqueId = (decltype(GetUniqueId))dlsym(h, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(h, "ncclCommInitRank");
        AllReduce = (decltype(AllReduce))dlsym(h, "ncclAllReduce");
        CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
        ok = GetUniqueId && CommInitRank && AllReduce && CommDestroy;
    }
};
inline NcclApi& nccl_api() {
    static NcclApi api;
    return api;
}

struct NcclCtx {
    ncclComm_t comm;
    int rank, nranks;
};

}  // namespace vb
