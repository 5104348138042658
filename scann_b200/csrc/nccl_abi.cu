// scann_allreduce_*: the NCCL gradient all-reduce of the data-parallel train step behind the C ABI (SURVEY.md section 8b,
// 8e: "NCCL comm init from a broadcast unique-id, flat-arena allreduce"; the reference itself trains on one device,
// scann/models/scann_model.py:232-241).  One communicator per process (one process per GPU): rank 0 calls
// scann_allreduce_unique_id, the 128 bytes travel to the other ranks by any host-side channel, every rank calls
// scann_allreduce_init with the device it will use current, and scann_allreduce_sum then sums the flat gradient arena
// (+ SSE) in place on the caller's stream -- it can be captured into the step's CUDA graph like the kernels around it.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy the process already has loaded -- e.g. the one PyTorch
// ships -- else the system one), so the library has no link-time dependency on it and loads on a box without NCCL.
#include <dlfcn.h>
#include <stddef.h>
#include <string.h>

#include "common.cuh"

namespace {
struct NcclId { char internal[128]; };                  // ncclUniqueId
typedef void* NcclComm;                                 // ncclComm_t
typedef int (*PFN_GetUniqueId)(NcclId*);
typedef int (*PFN_CommInitRank)(NcclComm*, int, NcclId, int);
typedef int (*PFN_AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*PFN_CommDestroy)(NcclComm);
typedef const char* (*PFN_GetErrorString)(int);
const int kNcclFloat32 = 7, kNcclSum = 0;               // nccl.h: ncclFloat32, ncclSum

struct Api {
    void* handle = nullptr;
    PFN_GetUniqueId get_id = nullptr;
    PFN_CommInitRank init_rank = nullptr;
    PFN_AllReduce all_reduce = nullptr;
    PFN_CommDestroy destroy = nullptr;
    PFN_GetErrorString err = nullptr;
} g_api;
NcclComm g_comm = nullptr;
int g_world = 0;

int load_api() {
    if (g_api.handle) return 0;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);           // already in the process (torch's copy)?
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { scann_set_error("scann_allreduce: libnccl.so.2 not found (%s)", dlerror()); return 1; }
    g_api.get_id = (PFN_GetUniqueId)dlsym(h, "ncclGetUniqueId");
    g_api.init_rank = (PFN_CommInitRank)dlsym(h, "ncclCommInitRank");
    g_api.all_reduce = (PFN_AllReduce)dlsym(h, "ncclAllReduce");
    g_api.destroy = (PFN_CommDestroy)dlsym(h, "ncclCommDestroy");
    g_api.err = (PFN_GetErrorString)dlsym(h, "ncclGetErrorString");
    if (!g_api.get_id || !g_api.init_rank || !g_api.all_reduce || !g_api.destroy) {
        scann_set_error("scann_allreduce: libnccl.so.2 lacks a required symbol");
        return 1;
    }
    g_api.handle = h;
    return 0;
}
int fail(const char* what, int rc) {
    scann_set_error("%s: NCCL error %d (%s)", what, rc, g_api.err ? g_api.err(rc) : "?");
    return 1;
}
}  // namespace

// out128: HOST buffer of 128 bytes (ncclUniqueId); call on ONE rank and send the bytes to the others.
extern "C" int scann_allreduce_unique_id(void* out128) {
    if (load_api()) return 1;
    NcclId id;
    const int rc = g_api.get_id(&id);
    if (rc != 0) return fail("scann_allreduce_unique_id", rc);
    memcpy(out128, &id, sizeof(id));
    return 0;
}

// Collective over all `world` ranks; the CUDA device that the rank will use must be current.
extern "C" int scann_allreduce_init(const void* id128, int rank, int world) {
    if (load_api()) return 1;
    if (g_comm) { scann_set_error("scann_allreduce_init: a communicator already exists (scann_allreduce_destroy first)"); return 1; }
    if (world < 1 || rank < 0 || rank >= world) { scann_set_error("scann_allreduce_init: bad rank %d of %d", rank, world); return 1; }
    NcclId id;
    memcpy(&id, id128, sizeof(id));
    const int rc = g_api.init_rank(&g_comm, world, id, rank);
    if (rc != 0) { g_comm = nullptr; return fail("scann_allreduce_init", rc); }
    g_world = world;
    return 0;
}

// buf[0..count) <- sum over the ranks, in place, fp32, on `stream` (asynchronous; capturable).
extern "C" int scann_allreduce_sum(float* buf, long long count, void* stream) {
    if (!g_comm) { scann_set_error("scann_allreduce_sum: no communicator (scann_allreduce_init)"); return 1; }
    if (count <= 0) return 0;
    const int rc = g_api.all_reduce(buf, buf, (size_t)count, kNcclFloat32, kNcclSum, g_comm, (cudaStream_t)stream);
    if (rc != 0) return fail("scann_allreduce_sum", rc);
    return 0;
}

extern "C" int scann_allreduce_world(void) { return g_comm ? g_world : 0; }

extern "C" int scann_allreduce_destroy(void) {
    if (!g_comm) return 0;
    const int rc = g_api.destroy(g_comm);
    g_comm = nullptr;
    g_world = 0;
    if (rc != 0) return fail("scann_allreduce_destroy", rc);
    return 0;
}
