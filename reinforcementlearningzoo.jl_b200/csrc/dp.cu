// Data-parallel communicator: one process per GPU, NCCL over NVLink 5 / NVSwitch.
// The reference has no distributed path at all; this is the §8(e) row: the learner update
// shards by sample, the only exchange is the gradient sum (+ the PER max(w) scalar), so one
// ncclAllReduce(sum) over the flat fp32 gradient buffer (with the loss riding in its tail) per update.
// NCCL is bound at run time with dlopen so the library has no link-time dependency on it and shares the
// copy a host process (e.g. torch) already loaded.
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

namespace zoo {

struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    // optional (NCCL >= 2.19): VMM-backed allocation + user-buffer registration, which is what lets the NVLS (in-switch) and
    // P2P algorithms read and write the gradient buffer in place instead of staging it through NCCL's own FIFO buffers
    ncclResult_t (*MemAlloc)(void **, size_t) = nullptr;
    ncclResult_t (*MemFree)(void *) = nullptr;
    ncclResult_t (*CommRegister)(const ncclComm_t, void *, size_t, void **) = nullptr;
    ncclResult_t (*CommDeregister)(const ncclComm_t, void *) = nullptr;
};
static NcclApi g_nccl;

static zoo_status load_nccl() {
    if (g_nccl.h) return ZOO_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        set_err("cannot dlopen libnccl.so.2: %s", dlerror());
        return ZOO_NCCL_ERROR;
    }
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
    g_nccl.MemAlloc = (decltype(g_nccl.MemAlloc))dlsym(h, "ncclMemAlloc");
    g_nccl.MemFree = (decltype(g_nccl.MemFree))dlsym(h, "ncclMemFree");
    g_nccl.CommRegister = (decltype(g_nccl.CommRegister))dlsym(h, "ncclCommRegister");
    g_nccl.CommDeregister = (decltype(g_nccl.CommDeregister))dlsym(h, "ncclCommDeregister");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce || !g_nccl.GetErrorString) {
        set_err("libnccl.so.2 lacks a required symbol");
        return ZOO_NCCL_ERROR;
    }
    g_nccl.h = h;
    return ZOO_OK;
}

#define ZOO_NCCL(call)                                                                              \
    do {                                                                                            \
        ncclResult_t r__ = (call);                                                                  \
        if (r__ != ncclSuccess) {                                                                   \
            zoo::set_err("%s:%d: %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r__)); \
            return ZOO_NCCL_ERROR;                                                                  \
        }                                                                                           \
    } while (0)

}  // namespace zoo

struct zoo_dp {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
};

namespace zoo {
int dp_world(const zoo_dp *dp) { return dp ? dp->world : 1; }
zoo_status dp_allreduce_sum(zoo_dp *dp, float *buf, long long n, cudaStream_t st) {
    if (!dp || dp->world == 1) return ZOO_OK;
    ZOO_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, ncclFloat, ncclSum, dp->comm, st));
    if (g_prof_on) { prof_label("allreduce[%lld]", n); prof_mark("ncclAllReduce", __LINE__); }
    return ZOO_OK;
}
// The flat gradient buffer of a data-parallel learner: allocated by NCCL (ncclMemAlloc: VMM-backed, multicast-capable) and
// registered with the communicator, so that all-reduces captured into the update's CUDA graph run zero-copy (NVLS / P2P direct)
// instead of copying through NCCL's staging buffers.  Falls back to the caller's cudaMalloc buffer when the NCCL in use has no
// such API (returns ZOO_OK with *ptr = nullptr).  ZOO_NCCL_REG=0 disables it (A/B runs).
zoo_status dp_alloc_registered(zoo_dp *dp, size_t bytes, void **ptr, void **handle) {
    *ptr = nullptr; *handle = nullptr;
    const char *e = getenv("ZOO_NCCL_REG");
    if (!dp || dp->world == 1 || (e && e[0] == '0') || !g_nccl.MemAlloc || !g_nccl.MemFree) return ZOO_OK;
    void *p = nullptr;
    if (g_nccl.MemAlloc(&p, bytes) != ncclSuccess || !p) { cudaGetLastError(); return ZOO_OK; }
    if (g_nccl.CommRegister && g_nccl.CommRegister(dp->comm, p, bytes, handle) != ncclSuccess) *handle = nullptr;
    *ptr = p;
    return ZOO_OK;
}
void dp_deregister(zoo_dp *dp, void *handle) {
    if (dp && dp->comm && handle && g_nccl.CommDeregister) g_nccl.CommDeregister(dp->comm, handle);
}
void dp_free_registered(void *ptr) {
    if (ptr && g_nccl.MemFree) g_nccl.MemFree(ptr);
}

zoo_status dp_allreduce_max(zoo_dp *dp, float *buf, long long n, cudaStream_t st) {
    if (!dp || dp->world == 1) return ZOO_OK;
    ZOO_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, ncclFloat, ncclMax, dp->comm, st));
    return ZOO_OK;
}
}  // namespace zoo

using namespace zoo;

extern "C" {

zoo_status zoo_dp_unique_id(uint8_t id_out[128]) {
    ZOO_REQUIRE(id_out, "zoo_dp_unique_id: null argument");
    ZOO_TRY(load_nccl());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    ZOO_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id_out, &id, 128);
    return ZOO_OK;
}

zoo_status zoo_dp_init(const uint8_t id[128], int32_t rank, int32_t world, zoo_dp **out) {
    ZOO_REQUIRE(id && out && world >= 1 && rank >= 0 && rank < world, "zoo_dp_init: bad argument");
    ZOO_TRY(require_sm100());
    ZOO_TRY(load_nccl());
    zoo_dp *dp = new zoo_dp();
    dp->rank = rank; dp->world = world;
    ncclUniqueId uid;
    memcpy(&uid, id, 128);
    ZOO_NCCL(g_nccl.CommInitRank(&dp->comm, world, uid, rank));
    if (world > 1) {
        const char *e = getenv("ZOO_DP_SM_BUDGET");
        if (e) g_sm_budget = std::max(32, std::min(148, atoi(e)));  // measured at N = 2: 148 / 128 / 112 within 1 %, so 148 stays
    }
    *out = dp;
    return ZOO_OK;
}

zoo_status zoo_dp_destroy(zoo_dp *dp) {
    if (!dp) return ZOO_OK;
    if (dp->comm) g_nccl.CommDestroy(dp->comm);
    delete dp;
    return ZOO_OK;
}

}  // extern "C"
