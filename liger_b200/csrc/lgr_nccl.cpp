// liger_b200 -- the only collective of the path: one sum all-reduce of the W/V normal equations per
// iteration.  NCCL over NVLink 5 / NVSwitch; symbols resolved with dlopen("libnccl.so.2").
#include "lgr_nccl.h"

#include <dlfcn.h>
#include <nccl.h>

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

struct NcclComm {
    ncclComm_t comm;
};

void nccl_unique_id(void* out128) {
    ncclUniqueId id;
    check(api().GetUniqueId(&id), "ncclGetUniqueId");
    memcpy(out128, &id, sizeof(id));
}
// Communicators are cached per process, keyed by (unique id, rank, world): ncclCommInitRank costs ~0.3 s, so a caller
// that passes the SAME unique id to successive lgr_inmf / lgr_uinmf / lgr_ctx_create calls (one per session, like an R
// session would) pays for it once.  A new id makes a new communicator.  Cached communicators live until process exit.
NcclComm* nccl_init(const void* id128, int rank, int world) {
    static std::mutex mu;
    static std::vector<std::pair<std::string, NcclComm*>> cache;
    std::string key(static_cast<const char*>(id128), sizeof(ncclUniqueId));
    key += ":" + std::to_string(rank) + "/" + std::to_string(world);
    std::lock_guard<std::mutex> lock(mu);
    for (auto& kv : cache)
        if (kv.first == key) return kv.second;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    NcclComm* c = new NcclComm;
    check(api().CommInitRank(&c->comm, world, id, rank), "ncclCommInitRank");
    cache.emplace_back(key, c);
    return c;
}
void nccl_allreduce_sum_f32(NcclComm* c, float* buf, size_t n, cudaStream_t st) {
    check(api().AllReduce(buf, buf, n, ncclFloat32, ncclSum, c->comm, st), "ncclAllReduce(f32)");
}
void nccl_allreduce_sum_f64(NcclComm* c, double* buf, size_t n, cudaStream_t st) {
    check(api().AllReduce(buf, buf, n, ncclFloat64, ncclSum, c->comm, st), "ncclAllReduce(f64)");
}
void nccl_group_start() { check(api().GroupStart(), "ncclGroupStart"); }
void nccl_group_end() { check(api().GroupEnd(), "ncclGroupEnd"); }
void nccl_destroy(NcclComm*) {}   // cached communicators are released at process exit
}  // namespace lgr
