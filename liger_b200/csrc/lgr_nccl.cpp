// liger_b200 -- the only collective of the path: one sum all-reduce of the W/V normal equations per
// iteration.  NCCL over NVLink 5 / NVSwitch; symbols resolved with dlopen("libnccl.so.2").
#include "lgr_nccl.h"

#include <dlfcn.h>
#include <nccl.h>

#include <memory>
#include <mutex>
#include <string>
#include <utility>
#include <vector>
#include <string.h>

#include "lgr_common.cuh"

namespace lgr {
namespace {
struct Api {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
Api& api() {
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            a.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (a.h) break;
        }
        if (!a.h) return;
#define LGR_SYM(field, name) a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.h, name))
        LGR_SYM(GetUniqueId, "ncclGetUniqueId");
        LGR_SYM(CommInitRank, "ncclCommInitRank");
        LGR_SYM(AllReduce, "ncclAllReduce");
        LGR_SYM(GroupStart, "ncclGroupStart");
        LGR_SYM(GroupEnd, "ncclGroupEnd");
        LGR_SYM(CommDestroy, "ncclCommDestroy");
        LGR_SYM(GetErrorString, "ncclGetErrorString");
#undef LGR_SYM
    });
    if (!a.h || !a.AllReduce || !a.CommInitRank || !a.GetUniqueId)
        throw Error(LGR_ERR_NCCL, "libnccl.so.2 could not be loaded (multi-GPU runs need NCCL)");
    return a;
}
void check(ncclResult_t r, const char* what) {
    if (r != ncclSuccess) throw Error(LGR_ERR_NCCL, std::string(what) + ": " + api().GetErrorString(r));
}
}  // namespace

// Communicators are cached per process, keyed by (unique id, rank, world, CUDA device): ncclCommInitRank costs ~0.3 s, so
// a caller that passes the SAME unique id to successive lgr_inmf / lgr_uinmf / lgr_ctx_create calls (one per session, like
// an R session would) pays for it once.  A new id makes a new communicator.  The cache holds one reference; contexts hold
// the others (nccl_destroy drops one).  At most MAX_CACHED idle communicators are kept: the oldest idle one is destroyed
// when a new id arrives, so a long-lived process that keeps creating ids does not accumulate them.
struct NcclComm {
    ncclComm_t comm = nullptr;
    int refs = 0;   // contexts using it
    std::string key;
};
namespace {
std::mutex g_mu;
std::vector<NcclComm*> g_cache;
constexpr size_t MAX_CACHED = 4;
}  // namespace

void nccl_unique_id(void* out128) {
    ncclUniqueId id;
    check(api().GetUniqueId(&id), "ncclGetUniqueId");
    memcpy(out128, &id, sizeof(id));
}
NcclComm* nccl_init(const void* id128, int rank, int world) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) throw Error(LGR_ERR_CUDA, "cudaGetDevice failed");
    std::string key(static_cast<const char*>(id128), sizeof(ncclUniqueId));
    key += ":" + std::to_string(rank) + "/" + std::to_string(world) + "@" + std::to_string(dev);
    std::lock_guard<std::mutex> lock(g_mu);
    for (NcclComm* c : g_cache)
        if (c->key == key) {
            c->refs += 1;
            return c;
        }
    // evict idle communicators beyond the cap (oldest first)
    for (size_t i = 0; i < g_cache.size() && g_cache.size() >= MAX_CACHED;) {
        if (g_cache[i]->refs == 0) {
            if (api().CommDestroy) api().CommDestroy(g_cache[i]->comm);
            delete g_cache[i];
            g_cache.erase(g_cache.begin() + i);
        } else {
            ++i;
        }
    }
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    std::unique_ptr<NcclComm> c(new NcclComm);   // not leaked when ncclCommInitRank fails
    check(api().CommInitRank(&c->comm, world, id, rank), "ncclCommInitRank");
    c->key = key;
    c->refs = 1;
    g_cache.push_back(c.get());
    return c.release();
}
void nccl_destroy(NcclComm* c) {   // the context lets go; the cache keeps the communicator for the next call with this id
    if (!c) return;
    std::lock_guard<std::mutex> lock(g_mu);
    if (c->refs > 0) c->refs -= 1;
}
void nccl_allreduce_sum_f32(NcclComm* c, float* buf, size_t n, cudaStream_t st) {
    check(api().AllReduce(buf, buf, n, ncclFloat32, ncclSum, c->comm, st), "ncclAllReduce(f32)");
}
void nccl_allreduce_sum_f64(NcclComm* c, double* buf, size_t n, cudaStream_t st) {
    check(api().AllReduce(buf, buf, n, ncclFloat64, ncclSum, c->comm, st), "ncclAllReduce(f64)");
}
void nccl_group_start() { check(api().GroupStart(), "ncclGroupStart"); }
void nccl_group_end() { check(api().GroupEnd(), "ncclGroupEnd"); }
}  // namespace lgr
