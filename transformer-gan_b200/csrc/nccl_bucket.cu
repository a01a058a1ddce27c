// Gradient all-reduce through the C-ABI: tgan_allreduce_bucket.  Replaces the implicit
// DistributedDataParallel all-reduce of train.py:649-655 (issued by autograd hooks inside loss.backward(), :904).
// The library does not link NCCL: it binds the process's libnccl.so.2 (the one torch ships and has already loaded)
// at run time, so the same .so works in single-GPU runs and on boxes without NCCL.  The caller owns the
// communicator (tgan_nccl_init from a broadcast unique id, one per process / GPU) and the stream: buckets are
// enqueued on a side stream while the backward of earlier layers is still running on the compute stream -- NCCL
// operations are stream-ordered and capturable, so the bucket launches also live inside the captured backward graph.
#include <dlfcn.h>
#include <string.h>

#include "common.cuh"

namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat32 = 7, ncclBfloat16 = 9, ncclSum = 0 };
struct Api {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
} g_api;

int load_api(const char* path) {
    if (g_api.AllReduce) return 0;
    const char* names[] = {path, "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n || !*n) continue;
        g_api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_api.handle) break;
    }
    if (!g_api.handle) {
        tgan_set_error("tgan_nccl: cannot load libnccl (%s)", dlerror());
        return 3;
    }
    g_api.GetUniqueId = (decltype(g_api.GetUniqueId))dlsym(g_api.handle, "ncclGetUniqueId");
    g_api.CommInitRank = (decltype(g_api.CommInitRank))dlsym(g_api.handle, "ncclCommInitRank");
    g_api.AllReduce = (decltype(g_api.AllReduce))dlsym(g_api.handle, "ncclAllReduce");
    g_api.CommDestroy = (decltype(g_api.CommDestroy))dlsym(g_api.handle, "ncclCommDestroy");
    g_api.GetErrorString = (decltype(g_api.GetErrorString))dlsym(g_api.handle, "ncclGetErrorString");
    if (!g_api.GetUniqueId || !g_api.CommInitRank || !g_api.AllReduce || !g_api.CommDestroy) {
        tgan_set_error("tgan_nccl: libnccl lacks a required symbol");
        g_api.AllReduce = nullptr;
        return 3;
    }
    return 0;
}
int check(ncclResult_t r, const char* what) {
    if (r == 0) return 0;
    tgan_set_error("%s failed: %s", what, g_api.GetErrorString ? g_api.GetErrorString(r) : "nccl error");
    return 2;
}
}  // namespace

/* lib_path: optional explicit path of libnccl.so.2 (NULL: the process's already-loaded / default one) */
extern "C" int tgan_nccl_load(const char* lib_path) { return load_api(lib_path); }

/* 128-byte unique id, created on one rank and distributed to the others by the caller (torch.distributed broadcast) */
extern "C" int tgan_nccl_unique_id(void* id128) {
    if (int rc = load_api(nullptr)) return rc;
    ncclUniqueId id;
    if (int rc = check(g_api.GetUniqueId(&id), "ncclGetUniqueId")) return rc;
    memcpy(id128, &id, sizeof(id));
    return 0;
}

/* collective: every rank calls it with the same id; the current CUDA device is the rank's GPU */
extern "C" int tgan_nccl_init(const void* id128, int nranks, int rank, void** comm_out) {
    if (int rc = load_api(nullptr)) return rc;
    TGAN_CHECK_ARG(id128 && comm_out && nranks >= 1 && rank >= 0 && rank < nranks, "tgan_nccl_init: bad arguments");
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    if (int rc = check(g_api.CommInitRank(&comm, nranks, id, rank), "ncclCommInitRank")) return rc;
    *comm_out = comm;
    return 0;
}

/* in-place SUM all-reduce of `count` elements (TGAN_F32 / TGAN_BF16) on `stream` */
extern "C" int tgan_allreduce_bucket(void* comm, void* buf, int64_t count, int dtype, void* stream) {
    TGAN_CHECK_ARG(comm && g_api.AllReduce, "tgan_allreduce_bucket: no communicator (call tgan_nccl_init first)");
    if (count <= 0) return 0;
    return check(g_api.AllReduce(buf, buf, (size_t)count, dtype == TGAN_F32 ? ncclFloat32 : ncclBfloat16, ncclSum,
                                 (ncclComm_t)comm, (cudaStream_t)stream),
                 "ncclAllReduce");
}

extern "C" int tgan_nccl_destroy(void* comm) {
    if (!comm || !g_api.CommDestroy) return 0;
    return check(g_api.CommDestroy((ncclComm_t)comm), "ncclCommDestroy");
}
