// The one collective of the data-parallel step behind the C ABI: an NCCL communicator owned by an
// opaque handle and ONE flat fp32 sum all-reduce over NVLink / NVSwitch per training step -- what replaces the
// reference's nn.DataParallel broadcast / gather / reduce-add (train_fine.py:122-123, train_coarse_fineFEAT.py:129-130).
//
// NCCL is resolved with dlopen("libnccl.so.2") at the first cf_comm_* call (in a PyTorch process this is the copy torch
// already loaded; libcfnet_b200.so has no link-time dependency on it).  The 128-byte unique id is created on one rank
// (cf_comm_unique_id) and handed to the others by the caller (any host channel: torch.distributed's store, a file, MPI).
#include "cf_common.cuh"
#include "../../include/cfnet_b200.h"
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

namespace {

typedef struct { char internal[128]; } cf_nccl_id;            // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef void* cf_nccl_comm;
enum { CF_NCCL_FLOAT = 7, CF_NCCL_SUM = 0 };                   // ncclFloat32, ncclSum (stable ABI values of nccl.h)

struct NcclApi {
    void* so = nullptr;
    int (*GetUniqueId)(cf_nccl_id*) = nullptr;
    int (*CommInitRank)(cf_nccl_comm*, int, cf_nccl_id, int) = nullptr;
    int (*CommDestroy)(cf_nccl_comm) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, cf_nccl_comm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static int state = 0;
    if (state == 0) {
        state = -1;
        api.so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!api.so) api.so = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (api.so) {
            *(void**)(&api.GetUniqueId) = dlsym(api.so, "ncclGetUniqueId");
            *(void**)(&api.CommInitRank) = dlsym(api.so, "ncclCommInitRank");
            *(void**)(&api.CommDestroy) = dlsym(api.so, "ncclCommDestroy");
            *(void**)(&api.AllReduce) = dlsym(api.so, "ncclAllReduce");
            *(void**)(&api.GetErrorString) = dlsym(api.so, "ncclGetErrorString");
            if (api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce) state = 1;
        }
    }
    return state == 1 ? &api : nullptr;
}

struct Comm {
    NcclApi* api;
    cf_nccl_comm comm;
    int world, rank;
};

int nccl_fail(NcclApi* api, const char* what, int rc) {
    cf_set_error("%s: NCCL error %d (%s)", what, rc, api->GetErrorString ? api->GetErrorString(rc) : "?");
    return CF_ERR_CUDA;
}

}  // namespace

extern "C" {

int cf_comm_unique_id(void* id128) {
    CF_CHECK_ARG(id128, "null pointer");
    NcclApi* api = nccl_api();
    if (!api) { cf_set_error("cf_comm_unique_id: libnccl.so.2 not found"); return CF_ERR_CUDA; }
    cf_nccl_id id;
    int rc = api->GetUniqueId(&id);
    if (rc != 0) return nccl_fail(api, "cf_comm_unique_id", rc);
    memcpy(id128, &id, sizeof(id));
    return CF_OK;
}

int cf_comm_init(void** comm, int world, int rank, const void* id128) {
    CF_CHECK_ARG(comm && id128 && world > 0 && rank >= 0 && rank < world, "bad argument");
    *comm = nullptr;
    NcclApi* api = nccl_api();
    if (!api) { cf_set_error("cf_comm_init: libnccl.so.2 not found"); return CF_ERR_CUDA; }
    Comm* c = (Comm*)calloc(1, sizeof(Comm));
    CF_CHECK_ARG(c, "out of host memory");
    c->api = api; c->world = world; c->rank = rank;
    cf_nccl_id id;
    memcpy(&id, id128, sizeof(id));
    int rc = api->CommInitRank(&c->comm, world, id, rank);          // the calling thread's current device joins
    if (rc != 0) { free(c); return nccl_fail(api, "cf_comm_init", rc); }
    *comm = c;
    return CF_OK;
}

int cf_comm_allreduce(void* comm, float* buf, int64_t n, cudaStream_t stream) {
    Comm* c = (Comm*)comm;
    CF_CHECK_ARG(c && buf && n > 0, "bad argument");
    int rc = c->api->AllReduce(buf, buf, (size_t)n, CF_NCCL_FLOAT, CF_NCCL_SUM, c->comm, stream);   // in place, fp32 sum
    if (rc != 0) return nccl_fail(c->api, "cf_comm_allreduce", rc);
    CF_COUNT_LAUNCH(1);
    return CF_OK;
}

int cf_comm_destroy(void* comm) {
    Comm* c = (Comm*)comm;
    if (!c) return CF_OK;
    c->api->CommDestroy(c->comm);
    free(c);
    return CF_OK;
}

}  // extern "C"
