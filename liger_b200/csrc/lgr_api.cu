// liger_b200 -- host side of the iNMF / UINMF path and the C ABI (include/liger_b200.h).
//
// The ANLS loop restates what RcppPlanc::inmf does behind rliger's .runINMF.list
// (reference R/integration.R:438-441): per iteration  H_i for all i  ->  V_i (U_i) for all i with the old W
// -> W with the new V_i   (update equations: reference R/cINMF.R:477-503), objective as src/objErr.cpp:13-31.
#include <string.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <unordered_map>

#include "lgr_common.cuh"
#include "lgr_nccl.h"

namespace lgr {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
cudaStream_t& alloc_stream() {
    static thread_local cudaStream_t s = nullptr;
    return s;
}

namespace {

struct KernelTimer {
    // CUDA events around every launch of the loop (on the launch stream), resolved after the loop.
    struct Rec {
        int id;
        cudaEvent_t a, b;
    };
    std::vector<cudaEvent_t> pool;
    std::vector<Rec> recs;
    size_t next = 0;
    bool on = false;
    std::vector<std::string> names;
    std::vector<double> ms;
    std::vector<int64_t> launches;
    int id_of(const char* name) {
        for (size_t i = 0; i < names.size(); ++i)
            if (names[i] == name) return (int)i;
        names.push_back(name);
        ms.push_back(0.0);
        launches.push_back(0);
        return (int)names.size() - 1;
    }
    cudaEvent_t get() {
        if (next == pool.size()) {
            cudaEvent_t e;
            LGR_CUDA(cudaEventCreate(&e));
            pool.push_back(e);
        }
        return pool[next++];
    }
    void begin_loop() {
        recs.clear();
        next = 0;
        std::fill(ms.begin(), ms.end(), 0.0);
        std::fill(launches.begin(), launches.end(), 0);
    }
    void resolve() {
        for (auto& r : recs) {
            float t = 0.f;
            LGR_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
            ms[r.id] += t;
            launches[r.id] += 1;
        }
        recs.clear();
        next = 0;
    }
    ~KernelTimer() {
        for (auto e : pool) cudaEventDestroy(e);
    }
};

struct Dataset {
    int n_full = 0, n = 0, c0 = 0, rows = 0, u = 0;
    size_t nnz = 0;
    double lambda = 0.0;
    DBuf<int> colptr, rowidx, tptr8, cptr8;
    DBuf<float> val, tval, cval, L, H;
    DBuf<unsigned short> tcell, crow;
    DBuf<unsigned int> ttag, ctag;
    int ngt = 1;
    DBuf<unsigned long long> hmask, vmask;
    DBuf<double> V, CtBv;
    size_t b_off = 0, r_off = 0;
    DenseLayout dl;   // dense-tile path (count-structured X)
};

}  // namespace
}  // namespace lgr

using namespace lgr;

// declared first in lgr_ctx => destroyed last, after every stream-ordered buffer has been freed on it
struct StreamHolder {
    cudaStream_t s = nullptr;
    ~StreamHolder() {
        if (s) {
            cudaStreamSynchronize(s);
            cudaStreamDestroy(s);
        }
    }
};

struct lgr_ctx {
    StreamHolder stream_holder;
    StreamHolder copy_holder;   // upload copy stream: buffers allocated in its order may live as long as the context
    StreamHolder side_holder;   // side stream of the iteration: the Gram inverses run beside the sweeps that do not need them
    cudaEvent_t ev_fork = nullptr, ev_join_h = nullptr, ev_join_v = nullptr;
    int d = 0, k = 0, kp = 32, m = 0, device = 0, rank = 0, world = 1, tc = 1024, cb = 512, gt = 1024, num_sms = 148;
    uint32_t flags = 0;
    int32_t init_seed = 0;
    cudaStream_t st = nullptr;
    std::vector<Dataset> ds;
    std::vector<DsDev> ds_host;
    DBuf<DsDev> ds_dev;
    DBuf<double> W, CtBw, Gsum, Garena, gram_arena, objbuf, sq;
    DBuf<unsigned long long> wmask;
    DBuf<float> Bbuf, Ubuf, Rarena;
    DBuf<NnlsGroup> grpT;      // tensor-core U = B P launch (tile offsets)
    long long tilesT = 0;
    bool dense = false;        // the two sweeps run as dense-tile tcgen05 contractions (lgr_dense.cu)
    int dense_mask = 3;        // developer knob LGR_DENSE_MASK: bit 0 = B sweep dense, bit 1 = R sweep dense
    std::vector<DenseDs> dds_host;   // one descriptor per (dataset, half of the factors): d for KP = 32, 2 d for KP = 64
    DBuf<DenseDs> dds_dev;
    int dds_n = 0, dense_max_n = 0;
    int k1_tiles = 0, gh_blocks = 0;
    long long k2_flat = 0;
    bool dshard = false;       // LGR_FLAG_DATASET_SHARD: this rank holds whole datasets; only the W solve's equations are all-reduced
    bool use_tc = false;
    bool fuse_u = false;       // U = B P in the tcgen05 epilogue of k_spmm_ltx_tiled (default on)
    DBuf<double> ctc_arena;   // per NNLS group: CtC and CtC^-1 (2 x KP x KP fp64), groups = H_0..H_d-1, V_0..V_d-1, W
    DBuf<int> ok_arena;
    DBuf<NnlsGroup> grpH, grpV, grpW;
    DBuf<GramSpec> specH, specV, specW;
    DBuf<Counters> counters;
    NnlsConfig cfgF{}, cfgD{};
    long long colsH = 0, colsV = 0; int ltx_blocks = 0, xh_tiles = 0, prep_blocks = 0, max_rows = 0;
    size_t N_local = 0, r_words = 0;
    NcclComm* comm = nullptr;
    bool stats_valid = false;  // R / G arenas hold the statistics of the current H
    bool h_valid = false;
    lgr_timings tm{};
    KernelTimer kt;   // per kernel
    KernelTimer pt;   // per phase: "h_phase" (Gram, B = L^T X, H solves), "vw_phase" (statistics, V and W solves), "comm"
    int (*interrupt)(void*) = nullptr;
    void* interrupt_user = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaGraphExec_t graph_exec = nullptr;   // one captured ANLS iteration
    int graph_cold = -1;
    int64_t graph_launches = 0;
    ~lgr_ctx() {
        if (graph_exec) cudaGraphExecDestroy(graph_exec);
        if (comm) nccl_destroy(comm);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join_h) cudaEventDestroy(ev_join_h);
        if (ev_join_v) cudaEventDestroy(ev_join_v);
    }
};

namespace {

// wall-clock trace of the host-side phases (LGR_TRACE=1)
struct Trace {
    bool on = getenv("LGR_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void mark(const char* what) {
        if (!on) return;
        cudaDeviceSynchronize();
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[lgr trace] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

struct Timed {
    lgr_ctx* c;
    int id = -1;
    cudaEvent_t a = nullptr, b = nullptr;
    Timed(lgr_ctx* c_, const char* name) : c(c_) {
        c->tm.kernel_launches += 1;
        if (c->kt.on) {
            id = c->kt.id_of(name);
            a = c->kt.get();
            b = c->kt.get();
            cudaEventRecord(a, c->st);
        }
    }
    ~Timed() {
        if (c->kt.on) {
            cudaEventRecord(b, c->st);
            c->kt.recs.push_back({id, a, b});
        }
    }
};

// CUDA events around a phase of the loop (only while kernel timing is on); resolved into lgr_timings after the loop
struct Span {
    lgr_ctx* c;
    int id = -1;
    cudaEvent_t a = nullptr, b = nullptr;
    Span(lgr_ctx* c_, const char* name) : c(c_) {
        if (c->kt.on) {
            id = c->pt.id_of(name);
            a = c->pt.get();
            b = c->pt.get();
            cudaEventRecord(a, c->st);
        }
    }
    ~Span() {
        if (c->kt.on) {
            cudaEventRecord(b, c->st);
            c->pt.recs.push_back({id, a, b});
        }
    }
};

// Device views handed out by lgr_prep_scaled_view: a view is only as alive as the buffers behind it.  Re-scaling a dataset
// (lgr_prep_scale) or destroying its pipeline frees them, so every view is registered by its column-pointer address and
// shape, forgotten when its buffers go, and checked before a factorisation dereferences it.
struct ViewInfo {
    int64_t nrow, ncol;
};
std::mutex g_view_mu;
std::unordered_map<const void*, ViewInfo> g_views;
void view_register(const void* colptr, int64_t nrow, int64_t ncol) {
    std::lock_guard<std::mutex> lk(g_view_mu);
    g_views[colptr] = ViewInfo{nrow, ncol};
}
void view_forget(const void* colptr) {
    if (!colptr) return;
    std::lock_guard<std::mutex> lk(g_view_mu);
    g_views.erase(colptr);
}
void view_require_live(const lgr_csc& M, const char* what) {
    std::lock_guard<std::mutex> lk(g_view_mu);
    auto it = g_views.find(M.colptr);
    if (it == g_views.end() || it->second.nrow != M.nrow || it->second.ncol != M.ncol)
        throw Error(LGR_ERR_INVALID, std::string(what) + ": stale or foreign device view (take it from lgr_prep_scaled_view AFTER the "
                                     "last lgr_prep_scale of that dataset, and keep the pipeline alive until the call returns)");
}

void check_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        throw Error(LGR_ERR_NO_DEVICE,
                    "liger_b200: no CUDA device is available; this library has no CPU path (the iNMF/UINMF hot path "
                    "runs only on the GPU)");
    }
}

// upload one CSC block's local column range [c0, c1) as fp32 device CSC
// Upload of one CSC block in two phases so that the PCIe copies of ALL blocks are queued on a copy stream up front
// and overlap the layout construction of the earlier blocks on the compute stream:
//   stage_csc   validates / rebases the column pointers on the host, allocates, enqueues the H2D copies on `cs`
//   finish_csc  (compute stream, after the block's copy event) converts fp64 values to fp32
struct CopyStream {
    cudaStream_t s = nullptr;
    CopyStream() { LGR_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)); }
    ~CopyStream() {
        if (s) {
            cudaStreamSynchronize(s);
            cudaStreamDestroy(s);
        }
    }
};
struct StagedCsc {
    int base = 0;             // first entry of the shard in the caller's arrays (column pointers are rebased on the device)
    DBuf<int> colptr, rowidx;
    DBuf<float> val;
    DBuf<double> val64, rs64, cs64;   // row / column scale of the count structure (dense path)
    size_t nnz = 0;
    cudaEvent_t done = nullptr;
    ~StagedCsc() {
        if (done) cudaEventDestroy(done);
    }
};
// Runs on the staging thread (its allocation stream is the copy stream, so the buffers need no cross-stream hand-over
// before the copies; the compute stream picks them up behind S.done).
void stage_csc(const lgr_csc& M, int c0, int c1, StagedCsc& S, cudaStream_t cs, bool pinned_promise, bool want_scales) {
    const int n = c1 - c0;
    const bool on_device = M.location == LGR_CSC_DEVICE;   // a view of resident data (lgr_prep_scaled_view)
    int base = 0, last = 0;
    if (on_device) {
        LGR_CUDA(cudaMemcpy(&base, M.colptr + c0, sizeof(int), cudaMemcpyDeviceToHost));
        LGR_CUDA(cudaMemcpy(&last, M.colptr + c0 + n, sizeof(int), cudaMemcpyDeviceToHost));
        LGR_REQUIRE(last >= base, "CSC column pointers must be non-decreasing");
    } else {
        base = M.colptr[c0];
        for (int j = 1; j <= n; ++j)
            LGR_REQUIRE(M.colptr[c0 + j] >= M.colptr[c0 + j - 1], "CSC column pointers must be non-decreasing");
        last = M.colptr[c0 + n];
    }
    LGR_REQUIRE(base >= 0, "CSC column pointers must be non-negative");
    S.base = base;
    S.nnz = (size_t)(last - base);
    S.colptr.alloc((size_t)n + 1);
    S.rowidx.alloc(S.nnz);
    S.val.alloc(S.nnz);
    if (S.nnz && M.value_dtype != LGR_F32) S.val64.alloc(S.nnz);
    LGR_CUDA(cudaEventCreateWithFlags(&S.done, cudaEventDisableTiming));
    const bool direct = pinned_promise || on_device;
    // straight from the caller's arrays when they are page-locked (or device views); through the bounce ring otherwise
    copy_in(S.colptr.p, M.colptr + c0, ((size_t)n + 1) * sizeof(int), cs, direct);
    if (S.nnz) {
        copy_in(S.rowidx.p, M.rowidx + base, S.nnz * sizeof(int), cs, direct);
        if (M.value_dtype == LGR_F32)
            copy_in(S.val.p, static_cast<const float*>(M.values) + base, S.nnz * sizeof(float), cs, direct);
        else
            copy_in(S.val64.p, static_cast<const double*>(M.values) + base, S.nnz * sizeof(double), cs, direct);
    }
    if (want_scales) {
        S.rs64.alloc((size_t)M.nrow);
        S.cs64.alloc((size_t)n);
        copy_in(S.rs64.p, M.row_scale, (size_t)M.nrow * sizeof(double), cs, direct);
        copy_in(S.cs64.p, M.col_scale + c0, (size_t)n * sizeof(double), cs, direct);
    }
    LGR_CUDA(cudaEventRecord(S.done, cs));
}
void finish_csc(StagedCsc& S, int nrow, int* bad, cudaStream_t st) {
    LGR_CUDA(cudaStreamWaitEvent(st, S.done, 0));
    if (S.base != 0) dev_add_int(S.colptr.p, S.colptr.n, -S.base, st);
    if (S.nnz && S.val64.p) dev_f64_to_f32(S.val64.p, S.val.p, S.nnz, st);
    // the tiled layouts index device arrays with the caller's row indices: range-check them first (one streaming pass)
    if (S.nnz) prep_check_rows(S.nnz, S.rowidx.p, nrow, bad, st);
}

// argument checks of one shared (E) or unshared (P) block; host pointers are only dereferenced for host matrices
void check_block(const lgr_csc& M, int64_t nrow, int64_t ncol, const char* what) {
    if (M.location == LGR_CSC_DEVICE) view_require_live(M, what);
    const std::string w(what);
    LGR_REQUIRE(M.nrow == nrow, w + ": wrong number of rows");
    LGR_REQUIRE(M.ncol == ncol, w + ": wrong number of cells");
    LGR_REQUIRE(M.ncol >= 1 && M.ncol < (int64_t)0x7fffffff, w + ": ncol out of range");
    LGR_REQUIRE(M.nrow >= 0 && M.nrow < (1 << 30), w + ": nrow out of range");
    LGR_REQUIRE(M.colptr, w + ": null column pointers");
    LGR_REQUIRE(M.value_dtype == LGR_F64 || M.value_dtype == LGR_F32, w + ": bad value_dtype");
    if (M.location == LGR_CSC_DEVICE) {
        LGR_REQUIRE(M.rowidx && M.values, w + ": null arrays");
    } else {
        LGR_REQUIRE(M.location == 0, w + ": bad location");
        LGR_REQUIRE(M.colptr[M.ncol] == 0 || (M.rowidx && M.values), w + ": null arrays");
    }
}

void upload_factor(const double* host, double* dev_padded, int rows, int k, int kp, cudaStream_t st) {
    if (rows == 0) return;
    DBuf<double> tmp((size_t)rows * k);
    copy_in(tmp.p, host, (size_t)rows * k * sizeof(double), st, false);
    dev_colmajor_to_padded(tmp.p, dev_padded, rows, k, kp, st);
    LGR_CUDA(cudaStreamSynchronize(st));
}

// Rows x k draws of R's runif(., 0, 2) (column-major, like matrix(runif(rows * k, 0, 2), rows, k)) into a padded factor
void draw_factor(lgr_rng* rng, double* dev_padded, int rows, int k, int kp, cudaStream_t st, std::vector<double>& tmp) {
    if (rows == 0) return;
    tmp.resize((size_t)rows * k);
    lgr_r_runif(rng, (int64_t)rows * k, 0.0, 2.0, tmp.data());
    upload_factor(tmp.data(), dev_padded, rows, k, kp, st);
}

void set_inits(lgr_ctx* c, const double* Winit, const double* const* Vinit, const double* const* Uinit) {
    // Winit == Vinit == NULL: RcppPlanc's internal random initialisation (R/suggestK.R:78-88 passes NULL for all three).
    // Drawn here from R's Mersenne-Twister stream in .runINMF.list's order (R/integration.R:412-437): W, V_1..V_d,
    // U_1..U_d; the H draws that follow in that order are only materialised when they are used (niter == 0).
    const bool draw = !Winit && !Vinit;
    LGR_REQUIRE(draw || (Winit && Vinit), "Winit and Vinit must both be given, or both be NULL (in-library seeded initialisation)");
    std::unique_ptr<lgr_rng, void (*)(lgr_rng*)> rng(nullptr, lgr_r_free);
    std::vector<double> tmp;
    if (draw) {
        rng.reset(lgr_r_set_seed(c->init_seed ? (uint32_t)c->init_seed : 1u));
        draw_factor(rng.get(), c->W.p, c->m, c->k, c->kp, c->st, tmp);
    } else {
        upload_factor(Winit, c->W.p, c->m, c->k, c->kp, c->st);
    }
    for (int i = 0; i < c->d; ++i) {
        Dataset& D = c->ds[i];
        if (draw) {
            draw_factor(rng.get(), D.V.p, c->m, c->k, c->kp, c->st, tmp);
        } else {
            LGR_REQUIRE(Vinit[i], "Vinit[i] is NULL");
            upload_factor(Vinit[i], D.V.p, c->m, c->k, c->kp, c->st);
        }
        D.hmask.zero(c->st);
        D.vmask.zero(c->st);
    }
    for (int i = 0; i < c->d; ++i) {
        Dataset& D = c->ds[i];
        if (D.u == 0) continue;
        if (draw) {
            draw_factor(rng.get(), D.V.p + (size_t)c->m * c->kp, D.u, c->k, c->kp, c->st, tmp);
        } else {
            LGR_REQUIRE(Uinit && Uinit[i], "Uinit[i] is required for datasets with unshared features");
            upload_factor(Uinit[i], D.V.p + (size_t)c->m * c->kp, D.u, c->k, c->kp, c->st);
        }
    }
    c->wmask.zero(c->st);
    c->counters.zero(c->st);
    c->stats_valid = false;
    c->h_valid = false;
    c->tm.nnls_pivots = c->tm.nnls_unconverged = 0;
}

void allreduce_stats(lgr_ctx* c) {
    if (c->world <= 1 || c->dshard) return;   // (dataset placement: the statistics of a dataset are complete on its rank)
    Span sp(c, "comm");   // includes the wait for the slowest rank
    // one collective: the fp64 Grams travel as (hi, lo) float pairs in the tail of the fp32 statistics arena
    const int ng = (int)c->Garena.n;
    float* tail = c->Rarena.p + c->r_words;
    launch_pack_gram(c->Garena.p, tail, ng, c->st);
    nccl_allreduce_sum_f32(c->comm, c->Rarena.p, c->Rarena.n, c->st);
    launch_unpack_gram(c->Garena.p, tail, ng, c->st);
    c->tm.kernel_launches += 2;
}

void run_prep(lgr_ctx* c) {
    c->gram_arena.zero(c->st);
    Timed t(c, "k_prep");
    launch_prep(c->ds_dev.p, c->d, c->prep_blocks, c->W.p, c->k, c->kp, c->st);
}
// the k x k Gram inverses are latency-bound single-CTA kernels: where the next big kernel does not read their result they
// run on the side stream (fork / join by events; inside a graph capture these become plain dependencies)
bool side_ok(lgr_ctx* c) { return c->side_holder.s != nullptr && !c->kt.on; }
void fork_gram_inv(lgr_ctx* c, const GramSpec* specs, int n, cudaEvent_t join) {
    LGR_CUDA(cudaEventRecord(c->ev_fork, c->st));
    LGR_CUDA(cudaStreamWaitEvent(c->side_holder.s, c->ev_fork, 0));
    launch_gram_inv(specs, n, c->k, c->kp, c->side_holder.s);
    LGR_CUDA(cudaEventRecord(join, c->side_holder.s));
    c->tm.kernel_launches += 1;
}

// R = X H, G = H^T H (+ all-reduce).  fork_ginv_v: also start the Gram inverses of the V solves as soon as G is final
// (single rank, dense path: right after k_gram_tc, beside the R sweep); returns whether it did.
bool run_stats(lgr_ctx* c, bool fork_ginv_v = false) {
    bool forked = false;
    c->Rarena.zero(c->st);
    c->Garena.zero(c->st);
    if (c->dense && (c->dense_mask & 2)) {
        {
            Timed t(c, "k_gram_tc");
            launch_gram_h(c->dds_dev.p, c->dds_n, c->gh_blocks, c->kp, c->num_sms, c->st);
            if (c->kp == 64) {   // the block of G between the two factor halves (the tensor-core kernel covers the diagonal blocks)
                launch_gram_cross(c->dds_dev.p, c->dds_n, c->dense_max_n, c->st);
                c->tm.kernel_launches += 1;
            }
        }
        if (fork_ginv_v && (c->world == 1 || c->dshard) && side_ok(c)) {
            fork_gram_inv(c, c->specV.p, c->d, c->ev_join_v);
            forked = true;
        }
        Timed t(c, "k_dense_xh");
        launch_dense_xh(c->dds_dev.p, c->dds_n, c->k2_flat, c->kp, c->num_sms, c->st);
    } else {
        Timed t(c, "k_spmm_xh");
        launch_spmm_xh(c->ds_dev.p, c->d, c->xh_tiles, c->k, c->kp, c->tc, c->st);
    }
    allreduce_stats(c);
    c->stats_valid = true;
    return forked;
}
void run_objective(lgr_ctx* c, double* slot) {
    LGR_CUDA(cudaMemsetAsync(slot, 0, sizeof(double), c->st));
    Timed t(c, "k_objective");
    launch_objective(c->ds_dev.p, c->d, c->k, c->kp, slot, c->st);
}

// FP64 solves (V_i / U_i, W): small batches go to the warp-per-column kernel, its singular groups to k_nnls
void solve_f64(lgr_ctx* c, const NnlsGroup* grp, int ngroups, long long cols, int flags) {
    static const bool no_warp = getenv("LGR_NNLS_NO_WARP") != nullptr;
    if (!no_warp && !(c->flags & LGR_FLAG_LEGACY_NNLS) && (flags & ~1) == 0 && nnls_warp_supported(c->k, c->kp, cols, c->num_sms)) {
        launch_nnls_warp(grp, ngroups, cols, c->k, c->num_sms, c->counters.p, flags, c->st);
        launch_nnls(true, grp, ngroups, cols, c->k, c->kp, c->cfgD, c->counters.p, flags | 16, c->st);
        c->tm.kernel_launches += 1;   // two kernels inside the caller's timed scope (which counts one)
    } else if (!no_warp && !(c->flags & LGR_FLAG_LEGACY_NNLS) && (flags & ~1) == 0 &&
               nnls_wbig64_supported(c->k, c->kp, cols, c->num_sms)) {
        launch_nnls_wbig64(grp, ngroups, cols, c->k, c->num_sms, c->counters.p, flags, c->st);
        launch_nnls(true, grp, ngroups, cols, c->k, c->kp, c->cfgD, c->counters.p, flags | 16, c->st);
        c->tm.kernel_launches += 1;
    } else {
        launch_nnls(true, grp, ngroups, cols, c->k, c->kp, c->cfgD, c->counters.p, flags, c->st);
    }
}

// The three block solves (R/cINMF.R:477-503), each from the factors currently on the device.  Everything is enqueued on
// c->st with no host synchronisation, so a sequence of them can be captured into a CUDA graph.
// H_i for every dataset:  CtC = L~^T L~ + lambda V~^T V~,  CtB = L~^T X  (obj_slot: objective of the iterate BEFORE the solve)
void enqueue_solve_h(lgr_ctx* c, int cold, double* obj_slot) {
    run_prep(c);
    if (obj_slot) run_objective(c, obj_slot);   // objective of the PREVIOUS iterate (uses the fresh L~ Grams)
    const bool fork = side_ok(c) && !c->fuse_u && !c->use_tc;   // (the fused U = B P epilogue of the gather K1 reads CtC^-1)
    if (fork) {
        fork_gram_inv(c, c->specH.p, c->d, c->ev_join_h);
    } else {
        Timed t(c, "k_gram_inv");   // CtC and CtC^-1 of the H solves
        launch_gram_inv(c->specH.p, c->d, c->k, c->kp, c->st);
    }
    if (c->dense && (c->dense_mask & 1)) {
        Timed t(c, "k_dense_ltx");
        launch_dense_ltx(c->dds_dev.p, c->dds_n, c->k1_tiles, c->kp, c->num_sms, c->st);
    } else {
        Timed t(c, "k_spmm_ltx");
        launch_spmm_ltx_tiled(c->ds_dev.p, c->d, c->ltx_blocks, c->kp, c->cb, c->gt, c->fuse_u, c->st);
    }
    if (c->use_tc) {
        Timed t(c, "k_pb_tc");
        launch_pb_tc(c->grpT.p, c->d, c->tilesT, c->kp, c->num_sms, c->st);
    }
    if (fork) LGR_CUDA(cudaStreamWaitEvent(c->st, c->ev_join_h, 0));
    {
        // fp32, k <= 32: register-resident solver; groups with a singular Gram fall through to k_nnls (flag 16)
        static const bool no_reg = getenv("LGR_NNLS_NO_REG") != nullptr;
        if (!no_reg && !(c->flags & LGR_FLAG_LEGACY_NNLS) && (cold & ~1) == 0 && nnls_reg_supported(c->k, c->kp)) {
            static const bool no_wbig = getenv("LGR_NNLS_NO_WBIG") != nullptr;
            {
                Timed t(c, "k_nnls_H");
                launch_nnls_reg(c->grpH.p, c->d, c->colsH, c->k, c->kp, c->num_sms, c->counters.p, cold, c->st);
            }
            if (c->kp > 32 && !no_wbig) {   // systems above 16 unknowns: continued from the passive set they had reached
                Timed t(c, "k_nnls_H_big");
                static const bool warp_cols = getenv("LGR_NNLS_WBIG") != nullptr;   // developer knob: the warp-per-column kernel
                if (warp_cols)
                    launch_nnls_wbig(c->grpH.p, c->d, c->colsH, c->k, c->num_sms, c->counters.p, c->st);
                else {
                    static const bool no_pair = getenv("LGR_NNLS_NO_PAIR") != nullptr;
                    if (!no_pair) launch_nnls_pair(c->grpH.p, c->d, c->colsH, c->k, c->num_sms, c->counters.p, c->st);
                    // systems above 26 unknowns (k > 52 only), or everything flagged without the pair kernel
                    if (no_pair || c->k > 52) launch_nnls_scol(c->grpH.p, c->d, c->colsH, c->k, c->num_sms, c->counters.p, c->st);
                }
            }
            Timed t(c, "k_nnls_H_mop");   // groups with a singular Gram (normally none: the kernel returns at once)
            launch_nnls(false, c->grpH.p, c->d, c->colsH, c->k, c->kp, c->cfgF, c->counters.p,
                        cold | 16 | ((c->kp > 32 && no_wbig) ? 64 : 0), c->st);
        } else {
            Timed t(c, "k_nnls_H");
            launch_nnls(false, c->grpH.p, c->d, c->colsH, c->k, c->kp, c->cfgF, c->counters.p, cold, c->st);
        }
    }
    c->h_valid = true;
    c->stats_valid = false;
}
// V_i (and U_i):  CtC = (1 + lambda) H^T H,  CtB = H^T X^T - H^T H W^T   (needs the statistics of the current H)
void enqueue_solve_v(lgr_ctx* c, int cold, bool ginv_forked = false) {
    {
        Timed t(c, "k_vrhs");
        launch_vrhs(c->ds_dev.p, c->d, c->W.p, c->k, c->kp, c->max_rows, c->st);
    }
    if (ginv_forked) {
        LGR_CUDA(cudaStreamWaitEvent(c->st, c->ev_join_v, 0));
    } else {
        Timed t(c, "k_gram_inv");
        launch_gram_inv(c->specV.p, c->d, c->k, c->kp, c->st);
    }
    {
        Timed t(c, "k_nnls_V");
        solve_f64(c, c->grpV.p, c->d, c->colsV, cold);
    }
}
// W:  CtC = sum_i H_i^T H_i,  CtB = sum_i (H_i^T X_i^T - H_i^T H_i V_i^T)
void enqueue_solve_w(lgr_ctx* c, int cold) {
    {
        Timed t(c, "k_wrhs");
        launch_wrhs(c->ds_dev.p, c->d, c->m, c->CtBw.p, c->Gsum.p, c->k, c->kp, c->st);
    }
    if (c->dshard) {   // the W solve's normal equations are sums over ALL datasets: the one collective of this placement
        Span sp(c, "comm");
        nccl_group_start();
        nccl_allreduce_sum_f64(c->comm, c->CtBw.p, c->CtBw.n, c->st);
        nccl_allreduce_sum_f64(c->comm, c->Gsum.p, c->Gsum.n, c->st);
        nccl_group_end();
    }
    {
        Timed t(c, "k_gram_inv");
        launch_gram_inv(c->specW.p, 1, c->k, c->kp, c->st);
    }
    {
        Timed t(c, "k_nnls_W");
        solve_f64(c, c->grpW.p, 1, (long long)c->m, cold);
    }
}

// One ANLS iteration: H_i for all i -> V_i (U_i) with the old W -> W with the new V_i.
void enqueue_iteration(lgr_ctx* c, int cold, double* obj_slot) {
    std::unique_ptr<Span> phase(new Span(c, "h_phase"));
    enqueue_solve_h(c, cold, obj_slot);
    phase.reset();
    phase.reset(new Span(c, "vw_phase"));
    const bool ginv_forked = run_stats(c, true);   // sufficient statistics of the V / W solves
    enqueue_solve_v(c, cold, ginv_forked);
    enqueue_solve_w(c, cold);
    phase.reset();
}

double objective(lgr_ctx* c);

void iterate(lgr_ctx* c, int niter, double* traj_host, double* final_obj = nullptr) {
    const bool traj = traj_host != nullptr;
    if (final_obj && niter == 0) {
        *final_obj = objective(c);
        return;
    }
    int cold = (c->flags & LGR_FLAG_COLD_START) ? 1 : 0;
    const bool hist = getenv("LGR_NNLS_HIST") != nullptr;
    if (hist) {
        cold |= 2;
        if (atoi(getenv("LGR_NNLS_HIST")) == 2) cold |= 4;  // histogram of final passive-set sizes instead
        c->counters.zero(c->st);
    }
    if (getenv("LGR_NNLS_NO_DUAL")) cold |= 8;
    if (traj) c->objbuf.alloc((size_t)niter + 1);
    c->kt.begin_loop();
    c->pt.begin_loop();
    c->tm.kernel_launches = 0;
    c->tm.h_phase_ms = c->tm.vw_phase_ms = c->tm.comm_ms = 0.0;
    LGR_CUDA(cudaEventRecord(c->ev0, c->st));
    // One iteration is ~17 launches + memsets (+ the NCCL all-reduce): replayed from a CUDA graph unless something needs
    // the host between kernels (per-kernel events, the trajectory's moving output slot, histograms).
    static const bool no_graph = getenv("LGR_NO_GRAPH") != nullptr;
    const bool use_graph = !no_graph && !c->kt.on && !traj && !hist && niter > 1;
    if (use_graph) {
        if (!c->graph_exec || c->graph_cold != cold) {
            if (c->graph_exec) {
                cudaGraphExecDestroy(c->graph_exec);
                c->graph_exec = nullptr;
            }
            const int64_t before = c->tm.kernel_launches;
            cudaGraph_t g = nullptr;
            LGR_CUDA(cudaStreamBeginCapture(c->st, cudaStreamCaptureModeThreadLocal));
            try {
                enqueue_iteration(c, cold, nullptr);
            } catch (...) {
                cudaStreamEndCapture(c->st, &g);
                if (g) cudaGraphDestroy(g);
                throw;
            }
            LGR_CUDA(cudaStreamEndCapture(c->st, &g));
            cudaError_t e = cudaGraphInstantiate(&c->graph_exec, g, 0);
            cudaGraphDestroy(g);
            LGR_CUDA(e);
            c->graph_cold = cold;
            c->graph_launches = c->tm.kernel_launches - before;
            c->tm.kernel_launches = before;
        }
        for (int it = 0; it < niter; ++it) {
            if (c->interrupt && c->interrupt(c->interrupt_user)) throw Error(LGR_ERR_INTERRUPT, "interrupted by caller");
            LGR_CUDA(cudaGraphLaunch(c->graph_exec, c->st));
            c->tm.kernel_launches += c->graph_launches;
            c->h_valid = true;
            c->stats_valid = true;
        }
    } else {
        for (int it = 0; it < niter; ++it) {
            if (c->interrupt && c->interrupt(c->interrupt_user)) throw Error(LGR_ERR_INTERRUPT, "interrupted by caller");
            enqueue_iteration(c, cold, (traj && it > 0) ? c->objbuf.p + (it - 1) : nullptr);
        }
    }
    DBuf<double> final_slot;
    if ((traj && niter > 0) || final_obj) {
        run_prep(c);
        if (traj && niter > 0) run_objective(c, c->objbuf.p + (niter - 1));
        if (final_obj) {
            final_slot.alloc(1);
            run_objective(c, final_slot.p);
        }
    }
    if (c->dshard) {   // every rank summed objErr_i over its own datasets
        if (traj && niter > 0) nccl_allreduce_sum_f64(c->comm, c->objbuf.p, (size_t)niter, c->st);
        if (final_obj) nccl_allreduce_sum_f64(c->comm, final_slot.p, 1, c->st);
    }
    LGR_CUDA(cudaEventRecord(c->ev1, c->st));
    LGR_CUDA(cudaStreamSynchronize(c->st));
    float ms = 0.f;
    LGR_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->tm.loop_ms = ms;
    if (final_obj) LGR_CUDA(cudaMemcpy(final_obj, final_slot.p, sizeof(double), cudaMemcpyDeviceToHost));
    if (c->kt.on) {
        c->kt.resolve();
        c->pt.resolve();
        for (size_t j = 0; j < c->pt.names.size(); ++j) {
            if (c->pt.names[j] == "h_phase") c->tm.h_phase_ms = c->pt.ms[j];
            if (c->pt.names[j] == "vw_phase") c->tm.vw_phase_ms = c->pt.ms[j];
            if (c->pt.names[j] == "comm") c->tm.comm_ms = c->pt.ms[j];
        }
    }
    if (traj && niter > 0)
        LGR_CUDA(cudaMemcpy(traj_host, c->objbuf.p, (size_t)niter * sizeof(double), cudaMemcpyDeviceToHost));
    Counters cn;
    LGR_CUDA(cudaMemcpy(&cn, c->counters.p, sizeof(cn), cudaMemcpyDeviceToHost));
    if (hist) {
        fprintf(stderr, "[lgr] nnls pivots/column histogram:");
        for (int b = 0; b < 40; ++b) fprintf(stderr, " %llu", cn.hist[b]);
        fprintf(stderr, "  spill=%llu bad=%llu\n", cn.nnls_spill, cn.nnls_bad_pivot);
    }
    c->tm.nnls_pivots = (int64_t)cn.nnls_pivots;
    c->tm.nnls_unconverged = (int64_t)cn.nnls_unconverged;
}

double objective(lgr_ctx* c) {
    if (!c->stats_valid) {
        LGR_REQUIRE(c->h_valid, "objective requested before any H is available (niter == 0 needs Hinit)");
        run_stats(c);
    }
    run_prep(c);
    if (c->objbuf.n < 1) c->objbuf.alloc(1);
    DBuf<double> slot(1);
    run_objective(c, slot.p);
    if (c->dshard) nccl_allreduce_sum_f64(c->comm, slot.p, 1, c->st);
    double v = 0.0;
    LGR_CUDA(cudaMemcpyAsync(&v, slot.p, sizeof(double), cudaMemcpyDeviceToHost, c->st));
    LGR_CUDA(cudaStreamSynchronize(c->st));
    return v;
}

// H_i (n_full x k column-major, R layout) -> this rank's rows [c0, c0 + n) of the padded fp32 H, passive sets from its signs
void upload_h(lgr_ctx* c, int i, const double* host) {
    Dataset& D = c->ds[i];
    if (D.n == 0) return;
    DBuf<double> tmp((size_t)D.n * c->k);
    LGR_CUDA(cudaMemcpy2DAsync(tmp.p, (size_t)D.n * sizeof(double), host + D.c0, (size_t)D.n_full * sizeof(double),
                               (size_t)D.n * sizeof(double), (size_t)c->k, cudaMemcpyHostToDevice, c->st));
    dev_colmajor_to_hpad(tmp.p, D.H.p, D.n, c->k, c->kp, c->st);
    dev_mask_from_positive(D.H.p, D.hmask.p, D.n, c->k, c->kp, c->st);
    LGR_CUDA(cudaStreamSynchronize(c->st));
}

// Replace any of the resident factors (NULL pointers / NULL entries keep what is on the device).
void set_factors(lgr_ctx* c, const double* W, const double* const* V, const double* const* U, const double* const* H) {
    if (W) upload_factor(W, c->W.p, c->m, c->k, c->kp, c->st);
    int nh = 0;
    for (int i = 0; i < c->d; ++i) {
        Dataset& D = c->ds[i];
        if (V && V[i]) upload_factor(V[i], D.V.p, c->m, c->k, c->kp, c->st);
        if (U && U[i] && D.u > 0) upload_factor(U[i], D.V.p + (size_t)c->m * c->kp, D.u, c->k, c->kp, c->st);
        if (H && H[i]) {
            upload_h(c, i, H[i]);
            ++nh;
        }
    }
    if (nh) {
        c->stats_valid = false;
        if (nh == c->d) c->h_valid = true;
    }
}

void read_counters(lgr_ctx* c) {
    Counters cn;
    LGR_CUDA(cudaMemcpy(&cn, c->counters.p, sizeof(cn), cudaMemcpyDeviceToHost));
    c->tm.nnls_pivots = (int64_t)cn.nnls_pivots;
    c->tm.nnls_unconverged = (int64_t)cn.nnls_unconverged;
}

// Block solves on the resident data, in the order H -> V -> W (the order of one ANLS iteration), each from the factors that
// are on the device when it starts.  The passive sets of the previous solve of the same block are the starting guess.
void solve_blocks(lgr_ctx* c, unsigned blocks) {
    LGR_REQUIRE(blocks != 0 && (blocks & ~(unsigned)(LGR_SOLVE_H | LGR_SOLVE_V | LGR_SOLVE_W)) == 0, "lgr_ctx_solve: unknown block");
    const int cold = (c->flags & LGR_FLAG_COLD_START) ? 1 : 0;
    if (blocks & LGR_SOLVE_H) enqueue_solve_h(c, cold, nullptr);
    if (blocks & (LGR_SOLVE_V | LGR_SOLVE_W)) {
        LGR_REQUIRE(c->h_valid, "lgr_ctx_solve: the V / W solves need an H (solve H first, or set it)");
        if (!c->stats_valid) run_stats(c);
        if (blocks & LGR_SOLVE_V) enqueue_solve_v(c, cold);
        if (blocks & LGR_SOLVE_W) enqueue_solve_w(c, cold);
    }
    LGR_CUDA(cudaStreamSynchronize(c->st));
    read_counters(c);
}

// The two products over X of the block solves, for callers that build their own normal equations
// (R/optimizeNewParam.R:97-134: residual right-hand sides):  which = 0:  out[i] = [X_i; P_i]^T ([W; 0] + [V_i; U_i])   n_i x k
//                                                             which = 1:  out[i] = [X_i; P_i] H_i                    (m + u_i) x k
void products(lgr_ctx* c, int which, double* const* out) {
    LGR_REQUIRE(out, "lgr_ctx_products: out is NULL");
    if (which == 0) {
        run_prep(c);
        if (c->dense && (c->dense_mask & 1))
            launch_dense_ltx(c->dds_dev.p, c->dds_n, c->k1_tiles, c->kp, c->num_sms, c->st);
        else
            launch_spmm_ltx_tiled(c->ds_dev.p, c->d, c->ltx_blocks, c->kp, c->cb, c->gt, false, c->st);
    } else {
        LGR_REQUIRE(c->h_valid, "lgr_ctx_products: X H needs an H (solve H first, or set it)");
        if (!c->stats_valid) run_stats(c);
    }
    for (int i = 0; i < c->d; ++i) {
        Dataset& D = c->ds[i];
        if (!out[i]) continue;
        const int rows = which == 0 ? D.n : D.rows;
        if (rows == 0) continue;
        const float* src = which == 0 ? c->Bbuf.p + D.b_off : c->Rarena.p + D.r_off;
        DBuf<double> tmp((size_t)rows * c->k);
        dev_hpad_to_colmajor(src, tmp.p, rows, c->k, c->kp, c->st);
        if (which == 0 && D.n != D.n_full) {   // slice mode: this rank's rows of the caller's n_full x k matrix
            LGR_CUDA(cudaMemcpy2DAsync(out[i] + D.c0, (size_t)D.n_full * sizeof(double), tmp.p, (size_t)D.n * sizeof(double),
                                       (size_t)D.n * sizeof(double), (size_t)c->k, cudaMemcpyDeviceToHost, c->st));
        } else {
            copy_out(out[i], tmp.p, (size_t)rows * c->k * sizeof(double), c->st, false);
        }
        LGR_CUDA(cudaStreamSynchronize(c->st));
    }
}

// objErr_i of ONE dataset from the resident factors (src/objErr.cpp:13-31); the sum over datasets is lgr_ctx_objective
double objective_i(lgr_ctx* c, int i) {
    LGR_REQUIRE(i >= 0 && i < c->d, "lgr_ctx_objerr_i: dataset index out of range");
    if (!c->stats_valid) {
        LGR_REQUIRE(c->h_valid, "objective requested before any H is available");
        run_stats(c);
    }
    run_prep(c);
    DBuf<double> slot(1);
    LGR_CUDA(cudaMemsetAsync(slot.p, 0, sizeof(double), c->st));
    launch_objective(c->ds_dev.p + i, 1, c->k, c->kp, slot.p, c->st);
    double v = 0.0;
    LGR_CUDA(cudaMemcpyAsync(&v, slot.p, sizeof(double), cudaMemcpyDeviceToHost, c->st));
    LGR_CUDA(cudaStreamSynchronize(c->st));
    return v;
}

void create(const lgr_inmf_args* a, lgr_ctx** out) {
    LGR_REQUIRE(a && out, "null argument");
    check_device();
    LGR_REQUIRE(a->d >= 1, "d must be >= 1");
    LGR_REQUIRE(a->k >= 1 && a->k <= LGR_MAX_K, "k must be in 1..64");
    LGR_REQUIRE(a->m >= 1 && a->m < (1 << 30), "m out of range");
    LGR_REQUIRE(a->E && a->lambda, "E and lambda are required");
    LGR_REQUIRE(a->niter >= 0, "niter must be >= 0");
    LGR_REQUIRE(a->world >= 1 && a->rank >= 0 && a->rank < a->world, "bad rank/world");
    Trace tr;
    std::unique_ptr<lgr_ctx> c(new lgr_ctx);
    c->d = a->d;
    c->k = a->k;
    c->kp = a->k <= 32 ? 32 : 64;
    c->m = (int)a->m;
    c->flags = a->flags;
    c->init_seed = a->init_seed;
    c->rank = a->rank;
    c->world = a->world;
    c->dshard = a->world > 1 && (a->flags & LGR_FLAG_DATASET_SHARD) != 0;
    if (c->dshard) LGR_REQUIRE(a->Winit && a->Vinit, "LGR_FLAG_DATASET_SHARD needs explicit Winit / Vinit (the seeded draw follows the global dataset order)");
    c->interrupt = a->interrupt;
    c->interrupt_user = a->interrupt_user;
    if (a->device >= 0) LGR_CUDA(cudaSetDevice(a->device));
    LGR_CUDA(cudaGetDevice(&c->device));
    cudaDeviceProp prop;
    LGR_CUDA(cudaGetDeviceProperties(&prop, c->device));
    LGR_REQUIRE(prop.major >= 10, "liger_b200 kernels are built for sm_100a (Blackwell B200) only");
    c->num_sms = prop.multiProcessorCount;
    kernels_init();
    LGR_CUDA(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
    c->stream_holder.s = c->st;
    AllocStreamScope alloc_scope(c->st);
    {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, c->device) == cudaSuccess) {
            unsigned long long keep = ~0ull;  // keep freed blocks cached for the next call
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    LGR_CUDA(cudaEventCreate(&c->ev0));
    LGR_CUDA(cudaEventCreate(&c->ev1));
    if (!getenv("LGR_NO_SIDE_STREAM")) {
        LGR_CUDA(cudaStreamCreateWithFlags(&c->side_holder.s, cudaStreamNonBlocking));
        LGR_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        LGR_CUDA(cudaEventCreateWithFlags(&c->ev_join_h, cudaEventDisableTiming));
        LGR_CUDA(cudaEventCreateWithFlags(&c->ev_join_v, cudaEventDisableTiming));
    }
    c->tc = c->kp == 32 ? 1024 : 512;
    c->cb = c->kp == 32 ? 512 : 256;
    {   // gene tiles of at most 1024 (512) rows, balanced
        const int gmax = c->kp == 32 ? 1024 : 512;
        int maxrows = (int)a->m;
        if (a->P)
            for (int i = 0; i < a->d; ++i) maxrows = std::max<int>(maxrows, (int)(a->m + a->P[i].nrow));
        const int nt = (maxrows + gmax - 1) / gmax;
        c->gt = (maxrows + nt - 1) / nt;
    }
    {   // Cell-block sizes of the two sweeps: one CTA per SM at a time, so the launch takes ceil(blocks / SMs) waves of a
        // duration ~ block size.  Pick the size (multiple of 16, within what shared memory allows) that minimises
        // waves x size over this rank's datasets: e.g. C3 (8 x 125 000 cells, 148 SMs) 1024 -> 1136 cells per K2 tile
        // turns 6.65 waves (7 x 1024) into exactly 6 (x 1136).
        const bool local = (c->flags & LGR_FLAG_LOCAL_SHARD) != 0 || c->dshard;
        std::vector<long long> ns;
        for (int i = 0; i < a->d; ++i) {
            const long long nf = a->E[i].ncol;
            ns.push_back(local ? nf : nf * (c->rank + 1) / c->world - nf * c->rank / c->world);
        }
        const size_t limit = (size_t)226 * 1024;
        auto pick = [&](int lo, int hi, auto fits) {
            int best = lo;
            double best_cost = 1e300;
            for (int v = lo; v <= hi; v += 16) {
                if (!fits(v)) break;
                long long blocks = 0;
                for (long long n : ns) blocks += (n + v - 1) / v;
                const long long waves = (blocks + c->num_sms - 1) / c->num_sms;
                const double cost = (double)waves * (v + 24);   // + per-block fixed cost (tile staging, epilogue)
                if (cost <= best_cost) {
                    best_cost = cost;
                    best = v;
                }
            }
            return best;
        };
        const int base_tc = c->tc, base_cb = c->cb;
        c->tc = pick(base_tc / 2, 2 * base_tc, [&](int v) { return spmm_xh_smem(c->kp, v) <= limit && (long long)v * (c->kp / 4) <= 65536; });
        if (!(c->flags & LGR_FLAG_TENSOR_U))   // the fused tensor-core epilogue needs cb x KP = 128 TMEM columns' worth
            c->cb = pick(base_cb / 2, 2 * base_cb, [&](int v) { return spmm_ltx_smem(c->kp, v, c->gt, false) <= limit; });
    }
    if (const char* e = getenv("LGR_CB")) {
        const int v = atoi(e);
        if (v >= 32 && v <= 1024 && v % 16 == 0) c->cb = v;
    }
    if (const char* e = getenv("LGR_TC")) {
        const int v = atoi(e);
        if (v >= 32 && v <= 65536 && spmm_xh_smem(c->kp, v) <= (size_t)226 * 1024) c->tc = v;
    }
    if (const char* e = getenv("LGR_KERNEL_TIMING")) c->kt.on = atoi(e) != 0;
    if (c->world > 1) {
        LGR_REQUIRE(a->nccl_unique_id, "world > 1 needs nccl_unique_id");
        c->comm = nccl_init(a->nccl_unique_id, c->rank, c->world);
    }
    cudaEvent_t u0, u1;
    LGR_CUDA(cudaEventCreate(&u0));
    LGR_CUDA(cudaEventCreate(&u1));
    LGR_CUDA(cudaEventRecord(u0, c->st));

    const int k = c->k, kp = c->kp, m = c->m;
    c->ds.resize(c->d);
    c->sq.alloc((size_t)c->d);
    c->sq.zero(c->st);
    size_t b_total = 0, r_total = 0;
    int max_rows = 0;
    tr.mark("ctx setup");
    // ---- pass 1: validate everything, then a staging thread moves the blocks to the device on a copy stream (direct DMA
    //      from page-locked buffers, through the bounce ring from pageable ones) while pass 2 builds the layouts ----
    const bool local = (c->flags & LGR_FLAG_LOCAL_SHARD) != 0 || c->dshard;
    const bool pinned_io = (c->flags & LGR_FLAG_PINNED_IO) != 0;
    for (int i = 0; i < c->d; ++i) {
        check_block(a->E[i], a->m, a->E[i].ncol, "E[i]");
        if (a->P && a->P[i].nrow > 0) check_block(a->P[i], a->P[i].nrow, a->E[i].ncol, "P[i]");
    }
    {   // dense-tile sweeps: every block (shared and unshared) carries a verified-on-upload count structure; k <= 32 in one
        // pass, 32 < k <= 64 in two passes over the entry streams (one per half of the factors)
        bool ok = (c->kp == 32 || c->kp == 64) && !(c->flags & (LGR_FLAG_NO_DENSE | LGR_FLAG_TENSOR_U | LGR_FLAG_LEGACY_NNLS)) && !getenv("LGR_NO_DENSE");
        for (int i = 0; i < c->d && ok; ++i) {
            ok = a->E[i].row_scale && a->E[i].col_scale;
            if (ok && a->P && a->P[i].nrow > 0) ok = a->P[i].row_scale && a->P[i].col_scale;
        }
        c->dense = ok;
        if (const char* e = getenv("LGR_DENSE_MASK")) c->dense_mask = atoi(e) & 3;
        if (c->dense_mask == 0) c->dense = false;
    }
    const bool dense = c->dense;
    // the copy stream belongs to the context: staged buffers are allocated (and later freed) in its stream order, and on an
    // error path they are released by the context's destructor
    struct { cudaStream_t s; } cstream{nullptr};
    LGR_CUDA(cudaStreamCreateWithFlags(&cstream.s, cudaStreamNonBlocking));
    c->copy_holder.s = cstream.s;
    std::vector<StagedCsc> stagedE(c->d), stagedP(c->d);
    struct StageSync {
        std::mutex mu;
        std::condition_variable cv;
        int staged = 0;   // datasets whose copies have been queued
        std::exception_ptr err;
    } sync;
    const int device = c->device;
    std::thread stager([&, device] {
        try {
            LGR_CUDA(cudaSetDevice(device));
            AllocStreamScope scope(cstream.s);
            for (int i = 0; i < c->d; ++i) {
                const int n_full = (int)a->E[i].ncol;
                const int s0 = local ? 0 : (int)((int64_t)n_full * c->rank / c->world);
                const int s1 = local ? n_full : (int)((int64_t)n_full * (c->rank + 1) / c->world);
                stage_csc(a->E[i], s0, s1, stagedE[i], cstream.s, pinned_io, dense);
                if (a->P && a->P[i].nrow > 0) stage_csc(a->P[i], s0, s1, stagedP[i], cstream.s, pinned_io, dense);
                {
                    std::lock_guard<std::mutex> g(sync.mu);
                    sync.staged = i + 1;
                }
                sync.cv.notify_all();
            }
        } catch (...) {
            std::lock_guard<std::mutex> g(sync.mu);
            sync.err = std::current_exception();
            sync.cv.notify_all();
        }
    });
    struct Joiner {
        std::thread& t;
        ~Joiner() {
            if (t.joinable()) t.join();
        }
    } joiner{stager};
    tr.mark("stage H2D copies (thread started)");
    DBuf<int> bad_rows(1);
    bad_rows.zero(c->st);
    // ---- pass 2: per dataset, wait for its copy and build the device layouts (the later copies keep streaming) ----
    for (int i = 0; i < c->d; ++i) {
        {
            std::unique_lock<std::mutex> g(sync.mu);
            sync.cv.wait(g, [&] { return sync.staged > i || sync.err; });
            if (sync.err) std::rethrow_exception(sync.err);
        }
        Dataset& D = c->ds[i];
        const lgr_csc& E = a->E[i];
        D.n_full = (int)E.ncol;
        D.c0 = local ? 0 : (int)((int64_t)D.n_full * c->rank / c->world);
        const int c1 = local ? D.n_full : (int)((int64_t)D.n_full * (c->rank + 1) / c->world);
        D.n = c1 - D.c0;
        D.lambda = a->lambda[i];
        D.u = 0;
        const lgr_csc* P = (a->P && a->P[i].nrow > 0) ? &a->P[i] : nullptr;
        if (P) D.u = (int)P->nrow;
        D.rows = m + D.u;
        max_rows = std::max(max_rows, D.rows);
        StagedCsc& SE = stagedE[i];
        StagedCsc& SP = stagedP[i];
        finish_csc(SE, m, bad_rows.p, c->st);
        if (P) finish_csc(SP, D.u, bad_rows.p, c->st);
        {
            int hbad = 0;
            LGR_CUDA(cudaMemcpyAsync(&hbad, bad_rows.p, sizeof(int), cudaMemcpyDeviceToHost, c->st));
            LGR_CUDA(cudaStreamSynchronize(c->st));
            LGR_REQUIRE(hbad == 0, "row index out of range in E[i] / P[i] (must be 0 <= i < nrow)");
        }
        if (!P) {
            D.colptr = std::move(SE.colptr);
            D.rowidx = std::move(SE.rowidx);
            D.val = std::move(SE.val);
            D.nnz = SE.nnz;
            tr.mark("finish_csc");
        } else {
            D.nnz = SE.nnz + SP.nnz;
            D.colptr.alloc((size_t)D.n + 1);
            D.rowidx.alloc(D.nnz);
            D.val.alloc(D.nnz);
            dev_stack_csc(D.n, SE.colptr.p, SE.rowidx.p, SE.val.p, SP.colptr.p, SP.rowidx.p, SP.val.p, m, D.colptr.p, D.rowidx.p,
                          D.val.p, c->st);
        }
        if (dense) {
            // count structure: X = N * rs * cs, verified entry by entry while the two entry streams are built
            D.dl.rs.alloc((size_t)D.rows);
            D.dl.cs.alloc((size_t)std::max(D.n, 1));
            dev_f64_to_f32(SE.rs64.p, D.dl.rs.p, (size_t)m, c->st);
            if (P) {   // stacked [E; P]: the unshared rows bring their own row scale, the cells' scale is shared
                dev_f64_to_f32(SP.rs64.p, D.dl.rs.p + m, (size_t)D.u, c->st);
                prep_check_same((size_t)D.n, SE.cs64.p, SP.cs64.p, bad_rows.p, c->st);
            }
            dev_f64_to_f32(SE.cs64.p, D.dl.cs.p, (size_t)D.n, c->st);
            dense_build(D.n, D.rows, D.colptr.p, D.rowidx.p, D.val.p, D.nnz, D.dl.rs.p, D.dl.cs.p, kp / 32, D.dl, bad_rows.p, c->st);
            int hbad = 0;
            LGR_CUDA(cudaMemcpyAsync(&hbad, bad_rows.p, sizeof(int), cudaMemcpyDeviceToHost, c->st));
            LGR_CUDA(cudaStreamSynchronize(c->st));
            LGR_REQUIRE(hbad == 0, "E[i] / P[i]: row_scale / col_scale do not factor the values into integer counts in 1..256, or "
                                   "the unshared block's col_scale differs from the shared block's (pass NULL scales, or "
                                   "LGR_FLAG_NO_DENSE, for general matrices)");
            tr.mark("build dense entry streams");
        }
        if (!dense || c->dense_mask != 3) {
            // the two padded tile formats of X (lgr_spmm.cu); the plain CSC upload is dropped afterwards
            dev_build_padded(1, D.n, D.rows, c->tc, kp / 4, 0, D.colptr.p, D.rowidx.p, D.val.p, D.nnz, D.tptr8, D.tcell, D.ttag, D.tval, c->st);
            tr.mark("build cell-tiled CSR");
            D.ngt = (D.rows + c->gt - 1) / c->gt;
            dev_build_padded(0, D.n, D.rows, c->gt, kp / 4, c->cb, D.colptr.p, D.rowidx.p, D.val.p, D.nnz, D.cptr8, D.crow, D.ctag, D.cval, c->st);
            tr.mark("build gene-tiled CSC");
        }
        dev_sumsq(D.val.p, D.nnz, c->sq.p + i, c->st);
        LGR_CUDA(cudaStreamSynchronize(c->st));
        D.colptr.release();
        D.rowidx.release();
        D.val.release();
        D.L.alloc((size_t)D.rows * kp);
        D.H.alloc((size_t)std::max(D.n, 1) * kp);
        D.H.zero(c->st);
        D.hmask.alloc((size_t)std::max(D.n, 1));
        D.vmask.alloc((size_t)D.rows);
        D.V.alloc((size_t)D.rows * kp);
        D.V.zero(c->st);
        D.CtBv.alloc((size_t)D.rows * kp);
        D.b_off = b_total;
        b_total += (size_t)D.n * kp;
        D.r_off = r_total;
        r_total += (size_t)D.rows * kp;
    }
    stager.join();
    LGR_CUDA(cudaStreamSynchronize(c->st));   // the staged buffers go back to the pool in the copy stream's order
    LGR_CUDA(cudaStreamSynchronize(cstream.s));
    stagedE.clear();
    stagedP.clear();
    c->max_rows = max_rows;
    c->N_local = b_total / kp;
    c->Bbuf.alloc(std::max<size_t>(b_total, 1));
    c->use_tc = (c->flags & LGR_FLAG_TENSOR_U) != 0;
    if (const char* e = getenv("LGR_TC_U")) c->use_tc = atoi(e) != 0;
    // LGR_FLAG_TENSOR_U: prefer the tcgen05 epilogue fused into k_spmm_ltx_tiled; the stand-alone k_pb_tc otherwise
    c->fuse_u = c->use_tc && (c->cb / 128) * c->kp == 128 && c->cb % 128 == 0 &&
                spmm_ltx_smem(c->kp, c->cb, c->gt, true) <= (size_t)227 * 1024 - 1024;
    if (const char* e = getenv("LGR_FUSE_U")) c->fuse_u = c->fuse_u && atoi(e) != 0;
    if (c->fuse_u) c->use_tc = false;
    if (c->use_tc || c->fuse_u) c->Ubuf.alloc(std::max<size_t>(b_total, 1));
    c->r_words = r_total;
    c->Rarena.alloc(r_total + (c->world > 1 ? (size_t)2 * c->d * kp * kp : 0));   // tail: the Grams as float pairs (all-reduce)
    c->Garena.alloc((size_t)c->d * kp * kp);
    c->gram_arena.alloc((size_t)2 * c->d * kp * kp);
    c->W.alloc((size_t)m * kp);
    c->CtBw.alloc((size_t)m * kp);
    c->Gsum.alloc((size_t)kp * kp);
    c->wmask.alloc((size_t)m);
    c->counters.alloc(1);
    if (c->world > 1 && !c->dshard) nccl_allreduce_sum_f64(c->comm, c->sq.p, (size_t)c->d, c->st);
    std::vector<double> sq((size_t)c->d);
    LGR_CUDA(cudaMemcpyAsync(sq.data(), c->sq.p, sq.size() * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    LGR_CUDA(cudaStreamSynchronize(c->st));

    // device-side dataset table + NNLS groups
    c->cfgF = nnls_config(k, kp, false, c->num_sms);
    c->cfgD = nnls_config(k, kp, true, c->num_sms);
    const int ngrp = 2 * c->d + 1;
    const size_t kk = (size_t)kp * kp;
    c->ctc_arena.alloc((size_t)ngrp * 2 * kk);
    c->ctc_arena.zero(c->st);
    c->ok_arena.alloc((size_t)ngrp);
    c->ok_arena.zero(c->st);
    auto ctc = [&](int gi) { return c->ctc_arena.p + (size_t)gi * 2 * kk; };
    std::vector<GramSpec> sH(c->d), sV(c->d), sW(1);
    c->ds_host.resize(c->d);
    std::vector<NnlsGroup> gH(c->d), gV(c->d), gW(1), gT(c->d);
    long long tT = 0;
    int ltx = 0, xh = 0, prep = 0;
    long long bH = 0, bV = 0;
    const int rb = prep_rows_per_block(kp);
    for (int i = 0; i < c->d; ++i) {
        Dataset& D = c->ds[i];
        DsDev& h = c->ds_host[i];
        memset(&h, 0, sizeof(h));
        h.n = D.n;
        h.rows = D.rows;
        h.m = m;
        h.u = D.u;
        h.ntiles = (D.n + c->tc - 1) / c->tc;
        h.ltx_blk_off = ltx;
        h.xh_blk_off = xh;
        h.prep_blk_off = prep;
        ltx += (D.n + c->cb - 1) / c->cb;
        xh += h.ntiles;
        prep += (D.rows + rb - 1) / rb;
        h.colptr = D.colptr.p;
        h.rowidx = D.rowidx.p;
        h.val = D.val.p;
        h.cptr8 = D.cptr8.p;
        h.crow = D.crow.p;
        h.ctag = D.ctag.p;
        h.cval = D.cval.p;
        h.tptr8 = D.tptr8.p;
        h.tcell = D.tcell.p;
        h.ttag = D.ttag.p;
        h.tval = D.tval.p;
        h.ngt = D.ngt;
        h.L = D.L.p;
        h.H = D.H.p;
        h.B = c->Bbuf.p + D.b_off;
        h.U = c->fuse_u ? c->Ubuf.p + D.b_off : nullptr;
        h.Pinv = ctc(i) + kk;
        h.hmask = D.hmask.p;
        h.V = D.V.p;
        h.CtBv = D.CtBv.p;
        h.vmask = D.vmask.p;
        h.R = c->Rarena.p + D.r_off;
        h.G = c->Garena.p + (size_t)i * kp * kp;
        h.LtL = c->gram_arena.p + (size_t)(2 * i) * kp * kp;
        h.VtV = c->gram_arena.p + (size_t)(2 * i + 1) * kp * kp;
        h.lambda = D.lambda;
        h.sqnorm = sq[i];
        h.Asplit = c->dense ? D.dl.Asplit.p : nullptr;
        h.rs = c->dense ? D.dl.rs.p : nullptr;
        h.asplit_half = c->dense ? D.dl.nstage1 * 8192 : 0;
        for (int half = 0; c->dense && half < kp / 32; ++half) {
            DenseDs q;
            memset(&q, 0, sizeof(q));
            q.n = D.n;
            q.rows = D.rows;
            q.ntile1 = D.dl.ntile1;
            q.nstage1 = D.dl.nstage1;
            q.k1_tile_off = c->k1_tiles;
            q.ngb = D.dl.ngb;
            q.ncstage = D.dl.ncstage;
            q.gh_blk_off = c->gh_blocks;
            q.k2_flat_off = c->k2_flat;
            q.seg1 = D.dl.seg1.p;
            q.ent1 = D.dl.ent1.p;
            q.ldf = kp;
            c->dense_max_n = std::max(c->dense_max_n, D.n);
            q.kloc = std::max(0, std::min(32, c->k - 32 * half));
            q.Asplit = D.dl.Asplit.p + (size_t)half * h.asplit_half;
            q.seg2 = D.dl.seg2.p;
            q.ent2 = D.dl.ent2.p;
            q.rs = D.dl.rs.p;
            q.cs = D.dl.cs.p;
            q.B = h.B + 32 * half;
            q.H = h.H + 32 * half;
            q.R = h.R + 32 * half;
            q.G = h.G + (size_t)(32 * half) * kp + 32 * half;
            q.Hs = D.dl.Hs.p + (size_t)half * ((size_t)D.dl.ncstage + 2) * 3072;
            c->k1_tiles += D.dl.ntile1;
            c->gh_blocks += (D.n + dense_gram_tile() - 1) / dense_gram_tile();
            c->k2_flat += (long long)D.dl.ngb * D.dl.ncstage;
            c->dds_host.push_back(q);
        }
        // H solve: CtC = LtL + lambda VtV, rhs = B (fp32), x = H
        sH[i] = GramSpec{h.LtL, h.VtV, 1.0, D.lambda, ctc(i), ctc(i) + kk, c->ok_arena.p + i};
        const float* upre = (c->use_tc || c->fuse_u) ? c->Ubuf.p + D.b_off : nullptr;
        gH[i] = NnlsGroup{ctc(i), ctc(i) + kk, c->ok_arena.p + i, h.B, upre, h.H, h.hmask, D.n, kp, bH};
        gT[i] = NnlsGroup{ctc(i), ctc(i) + kk, c->ok_arena.p + i, h.B, nullptr, c->use_tc ? (void*)(c->Ubuf.p + D.b_off) : nullptr,
                          nullptr, D.n, kp, tT};
        tT += (D.n + 127) / 128;
        bH += D.n;
        // V/U solve: CtC = (1 + lambda) G_i, rhs = CtBv (fp64), x = V
        sV[i] = GramSpec{h.G, nullptr, 1.0 + D.lambda, 0.0, ctc(c->d + i), ctc(c->d + i) + kk, c->ok_arena.p + c->d + i};
        gV[i] = NnlsGroup{ctc(c->d + i), ctc(c->d + i) + kk, c->ok_arena.p + c->d + i, h.CtBv, nullptr, h.V, h.vmask, D.rows, kp, bV};
        bV += D.rows;
    }
    sW[0] = GramSpec{c->Gsum.p, nullptr, 1.0, 0.0, ctc(2 * c->d), ctc(2 * c->d) + kk, c->ok_arena.p + 2 * c->d};
    gW[0] = NnlsGroup{ctc(2 * c->d), ctc(2 * c->d) + kk, c->ok_arena.p + 2 * c->d, c->CtBw.p, nullptr, c->W.p, c->wmask.p, m, kp, 0};
    c->ltx_blocks = ltx;
    c->xh_tiles = xh;
    c->prep_blocks = prep;
    c->colsH = bH;
    c->colsV = bV;
    if (c->dense) {
        c->dds_n = (int)c->dds_host.size();
        c->dds_dev.alloc((size_t)c->dds_n);
        LGR_CUDA(cudaMemcpyAsync(c->dds_dev.p, c->dds_host.data(), sizeof(DenseDs) * c->dds_n, cudaMemcpyHostToDevice, c->st));
    }
    c->ds_dev.alloc((size_t)c->d);
    c->grpH.alloc((size_t)c->d);
    c->grpV.alloc((size_t)c->d);
    c->grpW.alloc(1);
    c->grpT.alloc((size_t)c->d);
    c->tilesT = tT;
    LGR_CUDA(cudaMemcpyAsync(c->grpT.p, gT.data(), sizeof(NnlsGroup) * c->d, cudaMemcpyHostToDevice, c->st));
    c->specH.alloc((size_t)c->d);
    c->specV.alloc((size_t)c->d);
    c->specW.alloc(1);
    LGR_CUDA(cudaMemcpyAsync(c->specH.p, sH.data(), sizeof(GramSpec) * c->d, cudaMemcpyHostToDevice, c->st));
    LGR_CUDA(cudaMemcpyAsync(c->specV.p, sV.data(), sizeof(GramSpec) * c->d, cudaMemcpyHostToDevice, c->st));
    LGR_CUDA(cudaMemcpyAsync(c->specW.p, sW.data(), sizeof(GramSpec), cudaMemcpyHostToDevice, c->st));
    LGR_CUDA(cudaMemcpyAsync(c->ds_dev.p, c->ds_host.data(), sizeof(DsDev) * c->d, cudaMemcpyHostToDevice, c->st));
    LGR_CUDA(cudaMemcpyAsync(c->grpH.p, gH.data(), sizeof(NnlsGroup) * c->d, cudaMemcpyHostToDevice, c->st));
    LGR_CUDA(cudaMemcpyAsync(c->grpV.p, gV.data(), sizeof(NnlsGroup) * c->d, cudaMemcpyHostToDevice, c->st));
    LGR_CUDA(cudaMemcpyAsync(c->grpW.p, gW.data(), sizeof(NnlsGroup), cudaMemcpyHostToDevice, c->st));
    LGR_CUDA(cudaStreamSynchronize(c->st));

    tr.mark("tables + allocs");
    set_inits(c.get(), a->Winit, a->Vinit, a->Uinit);
    tr.mark("set_inits");
    if (a->Hinit) {
        for (int i = 0; i < c->d; ++i)
            if (a->Hinit[i]) upload_h(c.get(), i, a->Hinit[i]);
        c->h_valid = true;
    }
    LGR_CUDA(cudaEventRecord(u1, c->st));
    LGR_CUDA(cudaStreamSynchronize(c->st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, u0, u1);
    c->tm.upload_ms = ms;
    cudaEventDestroy(u0);
    cudaEventDestroy(u1);
    *out = c.release();
}

void download(lgr_ctx* c, lgr_inmf_out* o) {
    Trace tr;
    cudaEvent_t u0, u1;
    LGR_CUDA(cudaEventCreate(&u0));
    LGR_CUDA(cudaEventCreate(&u1));
    LGR_CUDA(cudaEventRecord(u0, c->st));
    const int k = c->k, kp = c->kp, m = c->m;
    const bool pinned_io = (c->flags & LGR_FLAG_PINNED_IO) != 0;
    // device layouts -> R layouts (column-major double) in scratch buffers, then one copy per output; pageable outputs are
    // drained through the bounce ring (lgr_stage.cu)
    auto factor_out = [&](const double* dev_padded, int rows, double* host) {
        if (!host || rows == 0) return;
        DBuf<double> tmp((size_t)rows * k);
        dev_padded_to_colmajor(dev_padded, tmp.p, rows, k, kp, c->st);
        copy_out(host, tmp.p, (size_t)rows * k * sizeof(double), c->st, pinned_io);
        LGR_CUDA(cudaStreamSynchronize(c->st));
    };
    factor_out(c->W.p, m, o->W);
    for (int i = 0; i < c->d; ++i) {
        Dataset& D = c->ds[i];
        if (o->V) factor_out(D.V.p, m, o->V[i]);
        if (o->U && D.u > 0) factor_out(D.V.p + (size_t)m * kp, D.u, o->U[i]);
        if (i == 0) tr.mark("download W,V");
        if (o->H && o->H[i] && D.n > 0) {
            DBuf<double> tmp((size_t)D.n * k);
            dev_hpad_to_colmajor(D.H.p, tmp.p, D.n, k, kp, c->st);
            if (D.n == D.n_full) {
                copy_out(o->H[i], tmp.p, (size_t)D.n * k * sizeof(double), c->st, pinned_io);
            } else {   // slice mode: this rank's rows [c0, c0 + n) of the caller's n_full x k matrix
                LGR_CUDA(cudaMemcpy2DAsync(o->H[i] + D.c0, (size_t)D.n_full * sizeof(double), tmp.p,
                                           (size_t)D.n * sizeof(double), (size_t)D.n * sizeof(double), (size_t)k,
                                           cudaMemcpyDeviceToHost, c->st));
            }
            LGR_CUDA(cudaStreamSynchronize(c->st));
        }
        if (o->H_passive && o->H_passive[i] && D.n > 0) {
            copy_out(o->H_passive[i] + D.c0, D.hmask.p, (size_t)D.n * sizeof(unsigned long long), c->st, pinned_io);
            LGR_CUDA(cudaStreamSynchronize(c->st));
        }
    }
    tr.mark("download rest (H)");
    LGR_CUDA(cudaEventRecord(u1, c->st));
    LGR_CUDA(cudaStreamSynchronize(c->st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, u0, u1);
    c->tm.download_ms = ms;
    cudaEventDestroy(u0);
    cudaEventDestroy(u1);
    o->timings = c->tm;
}

template <typename F>
int guarded(F&& f) {
    try {
        f();
        return LGR_OK;
    } catch (const Error& e) {
        set_last_error(e.what());
        return e.code;
    } catch (const std::bad_alloc&) {
        set_last_error("host allocation failed");
        return LGR_ERR_ALLOC;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return LGR_ERR_INVALID;
    } catch (...) {
        set_last_error("unknown error");
        return LGR_ERR_INVALID;
    }
}

int run_oneshot(const lgr_inmf_args* args, lgr_inmf_out* out, bool need_unshared) {
    return guarded([&] {
        LGR_REQUIRE(args && out, "null argument");
        if (need_unshared) {
            LGR_REQUIRE(args->P, "uinmf: unshared blocks (P) are required");
            bool any = false;
            for (int i = 0; i < args->d; ++i) any = any || args->P[i].nrow > 0;
            LGR_REQUIRE(any, "uinmf: at least one dataset must carry unshared features");
        }
        lgr_ctx* raw = nullptr;
        create(args, &raw);
        std::unique_ptr<lgr_ctx> c(raw);
        AllocStreamScope alloc_scope(c->st);
        iterate(c.get(), args->niter, (args->flags & LGR_FLAG_TRAJECTORY) ? out->obj_traj : nullptr, &out->objErr);
        download(c.get(), out);
    });
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

const char* lgr_last_error(void) { return g_last_error.c_str(); }

int lgr_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int lgr_version(void) { return LGR_VERSION_MAJOR * 1000 + LGR_VERSION_MINOR; }

int lgr_nccl_unique_id(void* out128) {
    return guarded([&] {
        LGR_REQUIRE(out128, "null argument");
        nccl_unique_id(out128);
    });
}

int lgr_inmf(const lgr_inmf_args* args, lgr_inmf_out* out) { return run_oneshot(args, out, false); }
int lgr_uinmf(const lgr_inmf_args* args, lgr_inmf_out* out) { return run_oneshot(args, out, true); }

int lgr_ctx_create(const lgr_inmf_args* args, lgr_ctx** ctx) {
    return guarded([&] { create(args, ctx); });
}
int lgr_ctx_reset(lgr_ctx* c, const double* Winit, const double* const* Vinit, const double* const* Uinit) {
    return guarded([&] {
        LGR_REQUIRE(c, "null ctx");
        LGR_CUDA(cudaSetDevice(c->device));
        AllocStreamScope alloc_scope(c->st);
        set_inits(c, Winit, Vinit, Uinit);
    });
}
int lgr_ctx_iterate(lgr_ctx* c, int32_t niter, double* obj_traj) {
    return guarded([&] {
        LGR_REQUIRE(c && niter >= 0, "bad argument");
        LGR_CUDA(cudaSetDevice(c->device));
        AllocStreamScope alloc_scope(c->st);
        iterate(c, niter, obj_traj);
    });
}
int lgr_ctx_run(lgr_ctx* c, int32_t niter, double* objErr) {
    return guarded([&] {
        LGR_REQUIRE(c && niter >= 0 && objErr, "bad argument");
        LGR_CUDA(cudaSetDevice(c->device));
        AllocStreamScope alloc_scope(c->st);
        iterate(c, niter, nullptr, objErr);
    });
}
int lgr_ctx_objective(lgr_ctx* c, double* obj) {
    return guarded([&] {
        LGR_REQUIRE(c && obj, "bad argument");
        LGR_CUDA(cudaSetDevice(c->device));
        AllocStreamScope alloc_scope(c->st);
        *obj = objective(c);
    });
}
int lgr_ctx_set_factors(lgr_ctx* c, const double* W, const double* const* V, const double* const* U, const double* const* H) {
    return guarded([&] {
        LGR_REQUIRE(c, "null ctx");
        LGR_CUDA(cudaSetDevice(c->device));
        AllocStreamScope alloc_scope(c->st);
        set_factors(c, W, V, U, H);
    });
}
int lgr_ctx_solve(lgr_ctx* c, uint32_t blocks) {
    return guarded([&] {
        LGR_REQUIRE(c, "null ctx");
        LGR_CUDA(cudaSetDevice(c->device));
        AllocStreamScope alloc_scope(c->st);
        solve_blocks(c, blocks);
    });
}
int lgr_ctx_products(lgr_ctx* c, int32_t which, double* const* out) {
    return guarded([&] {
        LGR_REQUIRE(c && (which == 0 || which == 1), "lgr_ctx_products: bad argument");
        LGR_CUDA(cudaSetDevice(c->device));
        AllocStreamScope alloc_scope(c->st);
        products(c, which, out);
    });
}
int lgr_ctx_objerr_i(lgr_ctx* c, int32_t i, double* obj) {
    return guarded([&] {
        LGR_REQUIRE(c && obj, "bad argument");
        LGR_CUDA(cudaSetDevice(c->device));
        AllocStreamScope alloc_scope(c->st);
        *obj = objective_i(c, i);
    });
}
// ---------------------------------------------------------------------------------------------------------------------
// online iNMF, scenarios 1-3 (kernels: lgr_online.cu; specification: DESIGN.md section 8)
// ---------------------------------------------------------------------------------------------------------------------
static int online_impl(const lgr_online_args* a, lgr_online_out* o, const lgr_chunk_source* src) {
    return guarded([&] {
        check_device();
        LGR_REQUIRE(!src || (src->read_cols && src->nnz), "lgr_online_inmf_chunked: the source needs read_cols and nnz");
        LGR_REQUIRE(a && o && a->E && o->H, "lgr_online_inmf: null argument");
        const bool project = (a->flags & LGR_ONLINE_PROJECT) != 0;   // scenario 3: H of the given datasets against Winit, nothing else
        const int nprev = a->n_prev;                                  // scenario 2: datasets [0, nprev) are already factorised
        LGR_REQUIRE(project || (o->W && o->V), "lgr_online_inmf: null argument");
        LGR_REQUIRE(nprev >= 0 && nprev < a->d, "online iNMF: n_prev must leave at least one dataset to learn from");
        LGR_REQUIRE(!project || (nprev == 0 && a->Winit), "online iNMF, projection: pass only the new datasets, and Winit");
        if (nprev > 0) {
            LGR_REQUIRE(a->Winit && a->Vinit && a->Ainit && a->Binit,
                        "Cannot find complete online iNMF result for current datasets (Winit, Vinit, Ainit, Binit)");
            for (int i = 0; i < nprev; ++i)
                LGR_REQUIRE(a->Vinit[i] && a->Ainit[i] && a->Binit[i],
                            "Cannot find complete online iNMF result for current datasets (Vinit[i], Ainit[i], Binit[i])");
        }
        LGR_REQUIRE(a->d >= 1 && a->d <= 16, "online iNMF: 1..16 datasets");
        LGR_REQUIRE(a->k >= 1 && a->k <= 32, "online iNMF: k must be in 1..32");
        LGR_REQUIRE(a->m >= 1 && a->m < (1 << 30), "m out of range");
        LGR_REQUIRE(a->max_epochs >= 1 && a->minibatch_size >= 1 && a->hals_iter >= 1, "online iNMF: bad iteration parameters");
        LGR_REQUIRE(project || nprev > 0 || (a->Winit == nullptr) == (a->Vinit == nullptr),
                    "Winit and Vinit must both be given, or both be NULL");
        const int d = a->d, k = a->k, kp = 32, m = (int)a->m;
        int ndev = 0;
        LGR_CUDA(cudaGetDeviceCount(&ndev));
        const int dev = a->device < 0 ? 0 : a->device;
        LGR_REQUIRE(dev < ndev, "device index out of range");
        LGR_CUDA(cudaSetDevice(dev));
        cudaDeviceProp prop;
        LGR_CUDA(cudaGetDeviceProperties(&prop, dev));
        StreamHolder sh;
        LGR_CUDA(cudaStreamCreateWithFlags(&sh.s, cudaStreamNonBlocking));
        cudaStream_t st = sh.s;
        AllocStreamScope alloc_scope(st);
        kernels_init();

        // ---- upload X_i (fp32 CSC), validate ----
        std::vector<int> n(d, 0);
        long long N = 0;
        std::vector<StagedCsc> S(d);
        DBuf<int> bad(1);
        bad.zero(st);
        // chunked source: every dataset is pulled through the caller's reader once, chunk by chunk, into the resident fp32 CSC
        // (the host never holds a whole matrix; HBM is the cache the reference's on-disk H5SpMat path does not have)
        const int64_t chunk_cols = src ? (src->chunk_cols > 0 ? src->chunk_cols : 1000) : 0;
        std::vector<int32_t> hcol, hrow;
        std::vector<double> hval;
        DBuf<double> v64;
        auto pull_chunk = [&](int i, int64_t c0, int64_t nc, size_t& cap) -> int64_t {   // fills hcol / hrow / hval, returns nnz
            for (;;) {
                hcol.assign((size_t)nc + 1, 0);
                if (hrow.size() < cap) {
                    hrow.resize(cap);
                    hval.resize(cap);
                }
                const int64_t got = src->read_cols(src->user, i, c0, nc, hcol.data(), hrow.data(), hval.data(), (int64_t)cap);
                if (got < 0 && (size_t)(-got) > cap) {   // the reader asks for more room
                    cap = (size_t)(-got);
                    continue;
                }
                LGR_REQUIRE(got >= 0 && (size_t)got <= cap, "chunk reader failed");
                LGR_REQUIRE(hcol[0] == 0 && hcol[(size_t)nc] == got, "chunk reader: column pointers must run from 0 to the chunk's nnz");
                for (int64_t j = 1; j <= nc; ++j) LGR_REQUIRE(hcol[(size_t)j] >= hcol[(size_t)j - 1], "chunk reader: column pointers must be non-decreasing");
                return got;
            }
        };
        size_t cap = src ? (size_t)std::max<int64_t>(src->max_chunk_nnz, 1024) : 0;
        for (int i = nprev; i < d; ++i) {
            if (!src) {
                check_block(a->E[i], m, a->E[i].ncol, "E[i]");
                LGR_REQUIRE(a->E[i].location == 0, "online iNMF takes host matrices");
            } else {
                LGR_REQUIRE(a->E[i].nrow == m && a->E[i].ncol >= 1 && a->E[i].ncol < (int64_t)0x7fffffff, "E[i]: bad shape");
                LGR_REQUIRE(src->nnz[i] >= 0 && src->nnz[i] < (int64_t)0x7fffffff, "E[i]: nnz out of range");
            }
            n[i] = (int)a->E[i].ncol;
            LGR_REQUIRE(n[i] >= k, "Number of factors (k) should be less than the number of cells in the smallest dataset.");
            N += n[i];
            if (!src) {
                stage_csc(a->E[i], 0, n[i], S[i], st, false, false);
                finish_csc(S[i], m, bad.p, st);
                continue;
            }
            StagedCsc& T = S[i];
            T.base = 0;
            T.nnz = (size_t)src->nnz[i];
            T.colptr.alloc((size_t)n[i] + 1);
            T.rowidx.alloc(std::max<size_t>(T.nnz, 1));
            T.val.alloc(std::max<size_t>(T.nnz, 1));
            int64_t running = 0;
            for (int64_t c0 = 0; c0 < n[i]; c0 += chunk_cols) {
                const int64_t nc = std::min<int64_t>(chunk_cols, n[i] - c0);
                const int64_t got = pull_chunk(i, c0, nc, cap);
                LGR_REQUIRE(running + got <= (int64_t)T.nnz, "chunk reader: more non-zeros than announced");
                for (int64_t j = 0; j <= nc; ++j) hcol[(size_t)j] += (int32_t)running;
                const bool last = c0 + nc == n[i];
                copy_in(T.colptr.p + c0, hcol.data(), (size_t)(nc + (last ? 1 : 0)) * sizeof(int), st, false);
                if (got) {
                    if (v64.n < (size_t)got) {
                        LGR_CUDA(cudaStreamSynchronize(st));
                        v64.release();
                        v64.alloc(std::max(cap, (size_t)got));
                    }
                    copy_in(T.rowidx.p + running, hrow.data(), (size_t)got * sizeof(int), st, false);
                    copy_in(v64.p, hval.data(), (size_t)got * sizeof(double), st, false);
                    dev_f64_to_f32(v64.p, T.val.p + running, (size_t)got, st);
                }
                running += got;
            }
            LGR_REQUIRE(running == (int64_t)T.nnz, "chunk reader: fewer non-zeros than announced");
            if (T.nnz) prep_check_rows(T.nnz, T.rowidx.p, m, bad.p, st);
        }
        LGR_REQUIRE(m >= k, "Number of factors (k) should be less than the number of shared features.");
        int hbad = 0;
        LGR_CUDA(cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        LGR_CUDA(cudaStreamSynchronize(st));
        LGR_REQUIRE(hbad == 0, "E[i]: row index out of range");

        // ---- initialisation (host): W ~ U(0, 2), V_i = k sampled cells of X_i, unit columns; then the minibatch schedule ----
        std::unique_ptr<lgr_rng, void (*)(lgr_rng*)> rng(lgr_r_set_seed(a->seed ? a->seed : 1u), lgr_r_free);
        RSampler rs{rng.get()};
        std::vector<int> scratch;
        std::vector<double> W0((size_t)m * k);
        std::vector<std::vector<double>> V0(d, std::vector<double>((size_t)m * k, 0.0));
        // (scenario 1 draws W; scenarios 2 / 3 start from the given W, and only the new datasets draw their V_i)
        if (nprev == 0 && !project) lgr_r_runif(rng.get(), (int64_t)m * k, 0.0, 2.0, W0.data());
        std::vector<int> picks(k);
        for (int i = nprev; i < d && !project; ++i) {
            rs.sample(n[i], k, scratch, picks.data());
            const lgr_csc& M = a->E[i];
            for (int j = 0; j < k; ++j) {
                const int c = picks[j];
                if (src) {   // the sampled cell, through the reader (in double precision, like the in-memory path)
                    const int64_t got = pull_chunk(i, c, 1, cap);
                    for (int64_t p = 0; p < got; ++p) {
                        LGR_REQUIRE(hrow[(size_t)p] >= 0 && hrow[(size_t)p] < m, "chunk reader: row index out of range");
                        V0[i][(size_t)j * m + hrow[(size_t)p]] += hval[(size_t)p];
                    }
                    continue;
                }
                for (int p = M.colptr[c]; p < M.colptr[c + 1]; ++p) {
                    const double v = M.value_dtype == LGR_F32 ? (double)static_cast<const float*>(M.values)[p]
                                                               : static_cast<const double*>(M.values)[p];
                    V0[i][(size_t)j * m + M.rowidx[p]] += v;
                }
            }
        }
        auto unit_columns = [&](std::vector<double>& X) {
            for (int j = 0; j < k; ++j) {
                double ss = 0.0;
                for (int g = 0; g < m; ++g) ss += X[(size_t)j * m + g] * X[(size_t)j * m + g];
                const double nrm = ss > 0.0 ? std::sqrt(ss) : 1.0;
                for (int g = 0; g < m; ++g) X[(size_t)j * m + g] /= nrm;
            }
        };
        if (nprev == 0 && !project) unit_columns(W0);
        for (int i = nprev; i < d && !project; ++i) unit_columns(V0[i]);
        // minibatch sizes and the walk of every dataset through its per-epoch permutations
        std::vector<int> mb(d), pos(d, 0);
        long long mb_total = 0;
        for (int i = nprev; i < d; ++i) {
            mb[i] = std::max(1, (int)std::nearbyint((double)n[i] / (double)N * (double)a->minibatch_size));
            mb[i] = std::min(mb[i], n[i]);
            mb_total += mb[i];
        }
        const long long iters = project ? 0 : (N * (long long)a->max_epochs) / a->minibatch_size;
        LGR_REQUIRE(project || iters >= 1, "online iNMF: minibatchSize is larger than maxEpochs x the number of cells");
        LGR_REQUIRE((double)iters * (double)mb_total < 2.0e9, "online iNMF: schedule too long");
        std::vector<std::vector<int>> perm(d);
        for (int i = nprev; i < d && !project; ++i) {
            perm[i].resize(n[i]);
            rs.sample(n[i], n[i], scratch, perm[i].data());
        }
        std::vector<int> sched((size_t)iters * mb_total);
        std::vector<long long> ds_off(d + 1, 0);
        for (int i = 0; i < d; ++i) ds_off[i + 1] = ds_off[i] + (i >= nprev ? mb[i] : 0);
        for (long long t = 0; t < iters; ++t) {
            for (int i = nprev; i < d; ++i) {
                int* dst = sched.data() + (size_t)t * mb_total + ds_off[i];
                int take = mb[i];
                while (take > 0) {
                    if (pos[i] == n[i]) {
                        rs.sample(n[i], n[i], scratch, perm[i].data());
                        pos[i] = 0;
                    }
                    const int c = std::min(take, n[i] - pos[i]);
                    memcpy(dst, perm[i].data() + pos[i], sizeof(int) * c);
                    dst += c;
                    pos[i] += c;
                    take -= c;
                }
            }
        }
        DBuf<int> d_sched(std::max<size_t>(sched.size(), 1));
        LGR_CUDA(cudaMemcpyAsync(d_sched.p, sched.data(), sizeof(int) * sched.size(), cudaMemcpyHostToDevice, st));

        // ---- device state ----
        DBuf<double> W((size_t)m * kp);
        std::vector<DBuf<double>> V(d), A(d), Bs(d);
        std::vector<DBuf<float>> Lt(d);
        DBuf<double> grams((size_t)d * 4 * kp * kp);   // per dataset: LtL, VtV, CtC, CtC^-1
        DBuf<int> ok(d);
        DBuf<float> Bmb((size_t)std::max<long long>(mb_total, 1) * kp), Hmb((size_t)std::max<long long>(mb_total, 1) * kp);
        DBuf<unsigned long long> masks((size_t)std::max<long long>(mb_total, 1));
        DBuf<Counters> counters(1);
        counters.zero(st);
        upload_factor(a->Winit ? a->Winit : W0.data(), W.p, m, k, kp, st);
        std::vector<const double*> hV(d);
        std::vector<float*> hLt(d);
        std::vector<double*> hLtL(d), hVtV(d);
        std::vector<GramSpec> specs(d);
        std::vector<NnlsGroup> grp(d);
        std::vector<OnlineDs> ods(d);
        std::vector<HalsDs> hds(d);
        for (int i = 0; i < d; ++i) {
            V[i].alloc((size_t)m * kp);
            A[i].alloc((size_t)kp * kp);
            Bs[i].alloc((size_t)m * kp);
            Lt[i].alloc((size_t)m * kp);
            A[i].zero(st);
            Bs[i].zero(st);
            if (i < nprev) {   // the state of an already factorised dataset (A_i is symmetric: either storage order)
                upload_factor(a->Ainit[i], A[i].p, k, k, kp, st);
                upload_factor(a->Binit[i], Bs[i].p, m, k, kp, st);
            }
            upload_factor((!project && a->Vinit && a->Vinit[i]) ? a->Vinit[i] : V0[i].data(), V[i].p, m, k, kp, st);
            hV[i] = V[i].p;
            hLt[i] = Lt[i].p;
            double* g4 = grams.p + (size_t)i * 4 * kp * kp;
            hLtL[i] = g4;
            hVtV[i] = g4 + kp * kp;
            specs[i] = GramSpec{hLtL[i], hVtV[i], 1.0, a->lambda, g4 + 2 * kp * kp, g4 + 3 * kp * kp, ok.p + i};
            grp[i] = NnlsGroup{g4 + 2 * kp * kp, g4 + 3 * kp * kp, ok.p + i, Bmb.p + (size_t)ds_off[i] * kp, nullptr,
                               Hmb.p + (size_t)ds_off[i] * kp, masks.p + ds_off[i], mb[i], kp, ds_off[i]};
            ods[i] = OnlineDs{Hmb.p + (size_t)ds_off[i] * kp, mb[i], A[i].p};
            hds[i] = HalsDs{V[i].p, A[i].p, Bs[i].p};
        }
        DBuf<const double*> dV(d);
        DBuf<float*> dLt(d);
        DBuf<double*> dLtL(d), dVtV(d);
        DBuf<GramSpec> dspecs(d);
        DBuf<NnlsGroup> dgrp(d);
        DBuf<OnlineDs> dods(d);
        DBuf<HalsDs> dhds(d);
        LGR_CUDA(cudaMemcpyAsync(dV.p, hV.data(), sizeof(void*) * d, cudaMemcpyHostToDevice, st));
        LGR_CUDA(cudaMemcpyAsync(dLt.p, hLt.data(), sizeof(void*) * d, cudaMemcpyHostToDevice, st));
        LGR_CUDA(cudaMemcpyAsync(dLtL.p, hLtL.data(), sizeof(void*) * d, cudaMemcpyHostToDevice, st));
        LGR_CUDA(cudaMemcpyAsync(dVtV.p, hVtV.data(), sizeof(void*) * d, cudaMemcpyHostToDevice, st));
        LGR_CUDA(cudaMemcpyAsync(dspecs.p, specs.data(), sizeof(GramSpec) * d, cudaMemcpyHostToDevice, st));
        LGR_CUDA(cudaMemcpyAsync(dgrp.p, grp.data(), sizeof(NnlsGroup) * d, cudaMemcpyHostToDevice, st));
        LGR_CUDA(cudaMemcpyAsync(dods.p, ods.data(), sizeof(OnlineDs) * d, cudaMemcpyHostToDevice, st));
        LGR_CUDA(cudaMemcpyAsync(dhds.p, hds.data(), sizeof(HalsDs) * d, cudaMemcpyHostToDevice, st));
        const int nnew = d - nprev;
        const NnlsConfig cfgF = nnls_config(k, kp, false, prop.multiProcessorCount);
        auto solve_h = [&](const NnlsGroup* g, int ng, long long cols) {   // every solve starts from the empty passive set
            launch_nnls_reg(g, ng, cols, k, kp, prop.multiProcessorCount, counters.p, 1, st);
            launch_nnls(false, g, ng, cols, k, kp, cfgF, counters.p, 1 | 16, st);
        };
        auto grams_now = [&]() {
            launch_online_gram(W.p, dV.p + nprev, nnew, m, k, dLt.p + nprev, dLtL.p + nprev, dVtV.p + nprev, st);
            launch_gram_inv(dspecs.p + nprev, nnew, k, kp, st);
        };
        cudaEvent_t e0, e1;
        LGR_CUDA(cudaEventCreate(&e0));
        LGR_CUDA(cudaEventCreate(&e1));
        LGR_CUDA(cudaEventRecord(e0, st));

        // ---- the minibatch loop: one iteration = ten launches driven by a device-side iteration counter, captured into a
        //      CUDA graph once and replayed (LGR_NO_GRAPH: launched one by one) ----
        std::vector<OnlineX> oxs(d);
        for (int i = nprev; i < d; ++i)
            oxs[i] = OnlineX{S[i].colptr.p, S[i].rowidx.p, S[i].val.p, Lt[i].p, Bs[i].p, 1.0 / (double)mb[i], (int)ds_off[i], 0};
        DBuf<OnlineX> doxs(d);
        DBuf<int> t_dev(1);
        const int one = 1;
        LGR_CUDA(cudaMemcpyAsync(doxs.p, oxs.data(), sizeof(OnlineX) * d, cudaMemcpyHostToDevice, st));
        LGR_CUDA(cudaMemcpyAsync(t_dev.p, &one, sizeof(int), cudaMemcpyHostToDevice, st));
        LGR_CUDA(cudaStreamSynchronize(st));
        auto one_iteration = [&]() {
            grams_now();
            launch_online_ltx_all(doxs.p + nprev, nnew, d_sched.p, (int)mb_total, t_dev.p, Bmb.p, st);
            solve_h(dgrp.p + nprev, nnew, mb_total);
            launch_online_stats_all(dods.p + nprev, doxs.p + nprev, nnew, m, k, d_sched.p, (int)mb_total, t_dev.p, Hmb.p, st);
            // W from the statistics of ALL datasets, V_i of the new ones only
            launch_online_hals(W.p, dhds.p, d, m, k, a->lambda, a->hals_iter, nprev, st);
            launch_online_next(t_dev.p, st);
        };
        if (project) {
            // scenario 3: nothing is learnt
        } else if (getenv("LGR_NO_GRAPH") || iters < 3) {
            for (long long t = 1; t <= iters; ++t) one_iteration();
        } else {
            one_iteration();   // un-captured first pass: lazy per-kernel initialisation (function attributes) happens here
            cudaGraph_t g = nullptr;
            cudaGraphExec_t ge = nullptr;
            LGR_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            try {
                one_iteration();
            } catch (...) {
                cudaStreamEndCapture(st, &g);
                if (g) cudaGraphDestroy(g);
                throw;
            }
            LGR_CUDA(cudaStreamEndCapture(st, &g));
            cudaError_t ce = cudaGraphInstantiate(&ge, g, 0);
            cudaGraphDestroy(g);
            LGR_CUDA(ce);
            for (long long t = 2; t <= iters; ++t) {
                ce = cudaGraphLaunch(ge, st);
                if (ce != cudaSuccess) break;
            }
            cudaGraphExecDestroy(ge);
            LGR_CUDA(ce);
        }

        // ---- final H over all cells, objective ----
        grams_now();
        std::vector<DBuf<float>> Hall(d), Ball(d);
        std::vector<DBuf<unsigned long long>> mall(d);
        std::vector<NnlsGroup> gall(d);
        long long off = 0;
        for (int i = nprev; i < d; ++i) {
            Hall[i].alloc((size_t)n[i] * kp);
            Ball[i].alloc((size_t)n[i] * kp);
            mall[i].alloc((size_t)n[i]);
            launch_online_ltx(S[i].colptr.p, S[i].rowidx.p, S[i].val.p, nullptr, n[i], Lt[i].p, Ball[i].p, st);
            const double* g4 = grams.p + (size_t)i * 4 * kp * kp;
            gall[i] = NnlsGroup{g4 + 2 * kp * kp, g4 + 3 * kp * kp, ok.p + i, Ball[i].p, nullptr, Hall[i].p, mall[i].p, n[i], kp, off};
            off += n[i];
        }
        DBuf<NnlsGroup> dgall(d);
        LGR_CUDA(cudaMemcpyAsync(dgall.p, gall.data(), sizeof(NnlsGroup) * d, cudaMemcpyHostToDevice, st));
        solve_h(dgall.p + nprev, nnew, off);
        // objErr = sum_i ||X_i||^2 - 2 <H_i, B_i> + tr(H_i^T H_i (L~^T L~ + lambda V^T V))
        DBuf<double> acc((size_t)2 * d), G((size_t)d * kp * kp);
        acc.zero(st);
        G.zero(st);
        std::vector<OnlineDs> gds(d);
        for (int i = nprev; i < d; ++i) {
            launch_online_sumsq(S[i].val.p, S[i].nnz, acc.p + 2 * i, st);
            launch_online_cross(Hall[i].p, Ball[i].p, (size_t)n[i] * kp, acc.p + 2 * i + 1, st);
            gds[i] = OnlineDs{Hall[i].p, n[i], G.p + (size_t)i * kp * kp};
        }
        DBuf<OnlineDs> dgds(d);
        LGR_CUDA(cudaMemcpyAsync(dgds.p, gds.data(), sizeof(OnlineDs) * d, cudaMemcpyHostToDevice, st));
        launch_online_stats_a(dgds.p + nprev, nnew, k, 0.0, st);   // H^T H / n_i
        LGR_CUDA(cudaEventRecord(e1, st));
        std::vector<double> hacc((size_t)2 * d), hG((size_t)d * kp * kp), hgr((size_t)d * 4 * kp * kp);
        LGR_CUDA(cudaMemcpyAsync(hacc.data(), acc.p, sizeof(double) * hacc.size(), cudaMemcpyDeviceToHost, st));
        LGR_CUDA(cudaMemcpyAsync(hG.data(), G.p, sizeof(double) * hG.size(), cudaMemcpyDeviceToHost, st));
        LGR_CUDA(cudaMemcpyAsync(hgr.data(), grams.p, sizeof(double) * hgr.size(), cudaMemcpyDeviceToHost, st));
        LGR_CUDA(cudaStreamSynchronize(st));
        double obj = 0.0;   // over the datasets that were read (scenario 2: the new ones)
        for (int i = nprev; i < d; ++i) {
            double tr = 0.0;
            const double* ctc = hgr.data() + (size_t)i * 4 * kp * kp + 2 * kp * kp;   // L~^T L~ + lambda V^T V
            for (int x = 0; x < k; ++x)
                for (int y = 0; y < k; ++y) {
                    double gxy = hG[(size_t)i * kp * kp + x * kp + y];
                    if (x == y && gxy == 1e-15) gxy = 0.0;   // (the lift of k_online_stats_a does not belong here)
                    tr += gxy * (double)n[i] * ctc[x * kp + y];
                }
            obj += hacc[2 * i] - 2.0 * hacc[2 * i + 1] + tr;
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        o->objErr = obj;
        o->iterations = (int32_t)iters;
        o->loop_ms = ms;

        // ---- results ----
        auto factor_out = [&](const double* dev_padded, int rows, double* host) {
            if (!host) return;
            DBuf<double> tmp((size_t)rows * k);
            dev_padded_to_colmajor(dev_padded, tmp.p, rows, k, kp, st);
            copy_out(host, tmp.p, (size_t)rows * k * sizeof(double), st, false);
            LGR_CUDA(cudaStreamSynchronize(st));
        };
        if (!project) factor_out(W.p, m, o->W);
        for (int i = 0; i < d; ++i) {
            if (!project) {
                if (o->V[i]) factor_out(V[i].p, m, o->V[i]);
                if (o->B && o->B[i]) factor_out(Bs[i].p, m, o->B[i]);
                if (o->A && o->A[i]) factor_out(A[i].p, k, o->A[i]);
            }
            if (i >= nprev && o->H[i]) {
                DBuf<double> tmp((size_t)n[i] * k);
                dev_hpad_to_colmajor(Hall[i].p, tmp.p, n[i], k, kp, st);
                copy_out(o->H[i], tmp.p, (size_t)n[i] * k * sizeof(double), st, false);
                LGR_CUDA(cudaStreamSynchronize(st));
            }
        }
    });
}

int lgr_online_inmf(const lgr_online_args* a, lgr_online_out* o) { return online_impl(a, o, nullptr); }
// the same over matrices that only exist as column chunks behind a reader (on-disk H5SpMat in the reference)
int lgr_online_inmf_chunked(const lgr_online_args* a, const lgr_chunk_source* src, lgr_online_out* o) {
    if (!src) return guarded([&] { LGR_REQUIRE(false, "lgr_online_inmf_chunked: null source"); });
    return online_impl(a, o, src);
}

// centroidAlign of host matrices
int lgr_centroid_align(int32_t d, int32_t k, const int64_t* n, const double* const* H, const lgr_ca_params* params,
                       double* const* aligned, int32_t* const* raw_which_max, int32_t* const* z_which_max,
                       int32_t* const* r_which_max, int32_t device) {
    return guarded([&] {
        check_device();
        LGR_REQUIRE(d >= 1 && n && H && params && aligned, "lgr_centroid_align: bad argument");
        int ndev = 0;
        LGR_CUDA(cudaGetDeviceCount(&ndev));
        const int dev = device < 0 ? 0 : device;
        LGR_REQUIRE(dev < ndev, "device index out of range");
        LGR_CUDA(cudaSetDevice(dev));
        StreamHolder sh;
        LGR_CUDA(cudaStreamCreateWithFlags(&sh.s, cudaStreamNonBlocking));
        AllocStreamScope alloc_scope(sh.s);
        const int kd = params->n_dims_use > 0 ? params->n_dims_use : k;
        std::vector<DBuf<double>> Hd(d), Od(d);
        std::vector<DBuf<int>> Dg(3 * (size_t)d);
        std::vector<const double*> hp(d);
        std::vector<double*> op(d);
        std::vector<int*> g0(d, nullptr), g1(d, nullptr), g2(d, nullptr);
        std::vector<int> nn(d);
        for (int i = 0; i < d; ++i) {
            LGR_REQUIRE(n[i] >= 0 && n[i] < (1LL << 31) && (n[i] == 0 || (H[i] && aligned[i])), "lgr_centroid_align: bad dataset");
            nn[i] = (int)n[i];
            Hd[i].alloc(std::max<size_t>((size_t)n[i] * k, 1));
            Od[i].alloc(std::max<size_t>((size_t)n[i] * kd, 1));
            if (n[i]) copy_in(Hd[i].p, H[i], (size_t)n[i] * k * sizeof(double), sh.s, false);
            hp[i] = Hd[i].p;
            op[i] = Od[i].p;
            if (raw_which_max && raw_which_max[i]) { Dg[3 * i].alloc(std::max<size_t>(n[i], 1)); g0[i] = Dg[3 * i].p; }
            if (z_which_max && z_which_max[i]) { Dg[3 * i + 1].alloc(std::max<size_t>(n[i], 1)); g1[i] = Dg[3 * i + 1].p; }
            if (r_which_max && r_which_max[i]) { Dg[3 * i + 2].alloc(std::max<size_t>(n[i], 1)); g2[i] = Dg[3 * i + 2].p; }
        }
        LGR_CUDA(cudaStreamSynchronize(sh.s));
        centroid_align_core(d, k, nn.data(), hp.data(), *params, op.data(), g0.data(), g1.data(), g2.data(), sh.s);
        for (int i = 0; i < d; ++i) {
            if (!n[i]) continue;
            copy_out(aligned[i], Od[i].p, (size_t)n[i] * kd * sizeof(double), sh.s, false);
            if (g0[i]) LGR_CUDA(cudaMemcpyAsync(raw_which_max[i], g0[i], (size_t)n[i] * sizeof(int), cudaMemcpyDeviceToHost, sh.s));
            if (g1[i]) LGR_CUDA(cudaMemcpyAsync(z_which_max[i], g1[i], (size_t)n[i] * sizeof(int), cudaMemcpyDeviceToHost, sh.s));
            if (g2[i]) LGR_CUDA(cudaMemcpyAsync(r_which_max[i], g2[i], (size_t)n[i] * sizeof(int), cudaMemcpyDeviceToHost, sh.s));
            LGR_CUDA(cudaStreamSynchronize(sh.s));
        }
    });
}
// centroidAlign of the resident H_i
int lgr_ctx_centroid_align(lgr_ctx* c, const lgr_ca_params* params, double* const* aligned) {
    return guarded([&] {
        LGR_REQUIRE(c && params && aligned, "lgr_ctx_centroid_align: bad argument");
        LGR_REQUIRE(c->world == 1, "lgr_ctx_centroid_align: the column statistics span all cells (single-rank contexts only)");
        LGR_REQUIRE(c->h_valid, "lgr_ctx_centroid_align: no H on the device yet");
        LGR_CUDA(cudaSetDevice(c->device));
        AllocStreamScope alloc_scope(c->st);
        const int d = c->d, k = c->k;
        const int kd = params->n_dims_use > 0 ? params->n_dims_use : k;
        std::vector<DBuf<double>> Hd(d), Od(d);
        std::vector<const double*> hp(d);
        std::vector<double*> op(d);
        std::vector<int> nn(d);
        for (int i = 0; i < d; ++i) {
            Dataset& D = c->ds[i];
            nn[i] = D.n;
            Hd[i].alloc(std::max<size_t>((size_t)D.n * k, 1));
            Od[i].alloc(std::max<size_t>((size_t)D.n * kd, 1));
            if (D.n) dev_hpad_to_colmajor(D.H.p, Hd[i].p, D.n, k, c->kp, c->st);
            hp[i] = Hd[i].p;
            op[i] = Od[i].p;
        }
        centroid_align_core(d, k, nn.data(), hp.data(), *params, op.data(), nullptr, nullptr, nullptr, c->st);
        for (int i = 0; i < d; ++i) {
            if (!nn[i] || !aligned[i]) continue;
            copy_out(aligned[i], Od[i].p, (size_t)nn[i] * kd * sizeof(double), c->st, false);
            LGR_CUDA(cudaStreamSynchronize(c->st));
        }
    });
}
// quantileNorm of host matrices: upload H_i (FP64), align on the device, download H.norm and the cluster labels
int lgr_quantile_norm(int32_t d, int32_t k, const int64_t* n, const double* const* H, const lgr_qn_params* params,
                      double* const* Hnorm, int32_t* const* clusters, lgr_qn_info* info, int32_t device) {
    return guarded([&] {
        check_device();
        LGR_REQUIRE(d >= 1 && n && H && params && Hnorm, "lgr_quantile_norm: bad argument");
        int ndev = 0;
        LGR_CUDA(cudaGetDeviceCount(&ndev));
        const int dev = device < 0 ? 0 : device;
        LGR_REQUIRE(dev < ndev, "device index out of range");
        LGR_CUDA(cudaSetDevice(dev));
        StreamHolder sh;
        LGR_CUDA(cudaStreamCreateWithFlags(&sh.s, cudaStreamNonBlocking));
        AllocStreamScope alloc_scope(sh.s);
        std::vector<DBuf<double>> Hd(d);
        std::vector<DBuf<int>> Ld(d);
        std::vector<double*> hp(d);
        std::vector<int*> lp(d);
        std::vector<int> nn(d);
        for (int i = 0; i < d; ++i) {
            LGR_REQUIRE(n[i] >= 1 && n[i] < (1LL << 31) && H[i] && Hnorm[i], "lgr_quantile_norm: bad dataset");
            nn[i] = (int)n[i];
            Hd[i].alloc((size_t)n[i] * k);
            Ld[i].alloc((size_t)n[i]);
            copy_in(Hd[i].p, H[i], (size_t)n[i] * k * sizeof(double), sh.s, false);
            hp[i] = Hd[i].p;
            lp[i] = Ld[i].p;
        }
        LGR_CUDA(cudaStreamSynchronize(sh.s));
        quantile_norm_core(d, k, nn.data(), hp.data(), *params, lp.data(), info, sh.s);
        for (int i = 0; i < d; ++i) {
            copy_out(Hnorm[i], Hd[i].p, (size_t)n[i] * k * sizeof(double), sh.s, false);
            if (clusters && clusters[i])
                LGR_CUDA(cudaMemcpyAsync(clusters[i], Ld[i].p, (size_t)n[i] * sizeof(int), cudaMemcpyDeviceToHost, sh.s));
            LGR_CUDA(cudaStreamSynchronize(sh.s));
        }
    });
}
// quantileNorm of the resident H_i: the fp32 H of the context is widened on the device; nothing is uploaded
int lgr_ctx_quantile_norm(lgr_ctx* c, const lgr_qn_params* params, double* const* Hnorm, int32_t* const* clusters,
                          lgr_qn_info* info) {
    return guarded([&] {
        LGR_REQUIRE(c && params && Hnorm, "lgr_ctx_quantile_norm: bad argument");
        LGR_REQUIRE(c->world == 1, "lgr_ctx_quantile_norm: the cells of a dataset must be on one GPU (single-rank contexts only)");
        LGR_REQUIRE(c->h_valid, "lgr_ctx_quantile_norm: no H on the device yet");
        LGR_CUDA(cudaSetDevice(c->device));
        AllocStreamScope alloc_scope(c->st);
        const int d = c->d, k = c->k;
        std::vector<DBuf<double>> Hd(d);
        std::vector<DBuf<int>> Ld(d);
        std::vector<double*> hp(d);
        std::vector<int*> lp(d);
        std::vector<int> nn(d);
        for (int i = 0; i < d; ++i) {
            Dataset& D = c->ds[i];
            nn[i] = D.n;
            Hd[i].alloc((size_t)D.n * k);
            Ld[i].alloc((size_t)D.n);
            dev_hpad_to_colmajor(D.H.p, Hd[i].p, D.n, k, c->kp, c->st);
            hp[i] = Hd[i].p;
            lp[i] = Ld[i].p;
        }
        quantile_norm_core(d, k, nn.data(), hp.data(), *params, lp.data(), info, c->st);
        for (int i = 0; i < d; ++i) {
            if (Hnorm[i]) copy_out(Hnorm[i], Hd[i].p, (size_t)nn[i] * k * sizeof(double), c->st, false);
            if (clusters && clusters[i])
                LGR_CUDA(cudaMemcpyAsync(clusters[i], Ld[i].p, (size_t)nn[i] * sizeof(int), cudaMemcpyDeviceToHost, c->st));
            LGR_CUDA(cudaStreamSynchronize(c->st));
        }
    });
}
int lgr_ctx_download(lgr_ctx* c, lgr_inmf_out* out) {
    return guarded([&] {
        LGR_REQUIRE(c && out, "bad argument");
        LGR_CUDA(cudaSetDevice(c->device));
        AllocStreamScope alloc_scope(c->st);
        download(c, out);
    });
}
int lgr_ctx_timings(lgr_ctx* c, lgr_timings* t) {
    return guarded([&] {
        LGR_REQUIRE(c && t, "bad argument");
        *t = c->tm;
    });
}
void* lgr_ctx_stream(lgr_ctx* c) { return c ? (void*)c->st : nullptr; }
int lgr_ctx_set_kernel_timing(lgr_ctx* c, int32_t on) {
    return guarded([&] {
        LGR_REQUIRE(c, "null ctx");
        c->kt.on = on != 0;
    });
}
int lgr_ctx_kernel_times(lgr_ctx* c, int32_t* n, const char** names, double* ms, int64_t* launches) {
    return guarded([&] {
        LGR_REQUIRE(c && n, "bad argument");
        const int cap = *n;
        const int cnt = (int)c->kt.names.size();
        for (int i = 0; i < cnt && i < cap; ++i) {
            if (names) names[i] = c->kt.names[i].c_str();
            if (ms) ms[i] = c->kt.ms[i];
            if (launches) launches[i] = c->kt.launches[i];
        }
        *n = cnt;
    });
}
void lgr_ctx_destroy(lgr_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    delete c;
}

void* lgr_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (bytes == 0) bytes = 1;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        set_last_error("cudaHostAlloc failed");
        return nullptr;
    }
    return p;
}
void lgr_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int lgr_bppnnls_prod(int32_t k, int64_t ncols, const double* CtC, const double* CtB, double* X, int32_t device) {
    return guarded([&] {
        check_device();
        LGR_REQUIRE(k >= 1 && k <= LGR_MAX_K, "k must be in 1..64");
        LGR_REQUIRE(ncols >= 0 && ncols < (int64_t)0x7fffffff, "column count out of range");
        LGR_REQUIRE(CtC && (ncols == 0 || (CtB && X)), "null argument");
        if (ncols == 0) return;
        if (device >= 0) LGR_CUDA(cudaSetDevice(device));
        kernels_init();
        cudaDeviceProp prop;
        int dev = 0;
        LGR_CUDA(cudaGetDevice(&dev));
        LGR_CUDA(cudaGetDeviceProperties(&prop, dev));
        const int kp = k <= 32 ? 32 : 64;
        cudaStream_t st;
        LGR_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        AllocStreamScope alloc_scope(st);
        try {
            DBuf<double> dG((size_t)k * k), dGp((size_t)kp * kp), dB((size_t)k * ncols), dX((size_t)k * ncols);
            DBuf<unsigned long long> mask((size_t)ncols);
            DBuf<Counters> cnt(1);
            DBuf<NnlsGroup> grp(1);
            mask.zero(st);
            cnt.zero(st);
            LGR_CUDA(cudaMemcpyAsync(dG.p, CtC, (size_t)k * k * sizeof(double), cudaMemcpyHostToDevice, st));
            LGR_CUDA(cudaMemcpyAsync(dB.p, CtB, (size_t)k * ncols * sizeof(double), cudaMemcpyHostToDevice, st));
            dev_colmajor_to_padded(dG.p, dGp.p, k, k, kp, st);  // symmetric: layout orientation is irrelevant
            NnlsConfig cfg = nnls_config(k, kp, true, prop.multiProcessorCount);
            DBuf<double> ctc((size_t)2 * kp * kp);
            DBuf<int> ok(1);
            DBuf<GramSpec> spec(1);
            GramSpec sp{dGp.p, nullptr, 1.0, 0.0, ctc.p, ctc.p + (size_t)kp * kp, ok.p};
            LGR_CUDA(cudaMemcpyAsync(spec.p, &sp, sizeof(sp), cudaMemcpyHostToDevice, st));
            launch_gram_inv(spec.p, 1, k, kp, st);
            NnlsGroup g{ctc.p, ctc.p + (size_t)kp * kp, ok.p, dB.p, nullptr, dX.p, mask.p, (int)ncols, k, 0};
            LGR_CUDA(cudaMemcpyAsync(grp.p, &g, sizeof(g), cudaMemcpyHostToDevice, st));
            static const bool no_warp = getenv("LGR_NNLS_NO_WARP") != nullptr;
            if (!no_warp && nnls_warp_supported(k, kp, ncols, prop.multiProcessorCount)) {
                launch_nnls_warp(grp.p, 1, (long long)ncols, k, prop.multiProcessorCount, cnt.p, 1, st);
                launch_nnls(true, grp.p, 1, (long long)ncols, k, kp, cfg, cnt.p, 1 | 16, st);
            } else {
                launch_nnls(true, grp.p, 1, (long long)ncols, k, kp, cfg, cnt.p, 1, st);
            }
            LGR_CUDA(cudaMemcpyAsync(X, dX.p, (size_t)k * ncols * sizeof(double), cudaMemcpyDeviceToHost, st));
            LGR_CUDA(cudaStreamSynchronize(st));
        } catch (...) {
            cudaStreamDestroy(st);
            throw;
        }
        cudaStreamDestroy(st);
    });
}

int lgr_objerr_i(int32_t k, const double* H, const double* W, const double* V, const lgr_csc* E, double lambda,
                 int32_t device, double* err) {
    return guarded([&] {
        LGR_REQUIRE(H && W && V && E && err, "null argument");
        // H arrives k x n (src/objErr.cpp:13); the context wants n x k column-major
        const int64_t n = E->ncol;
        std::vector<double> Ht((size_t)n * k);
        for (int64_t c = 0; c < n; ++c)
            for (int f = 0; f < k; ++f) Ht[(size_t)f * n + c] = H[(size_t)c * k + f];
        lgr_inmf_args a;
        memset(&a, 0, sizeof(a));
        a.d = 1;
        a.k = k;
        a.niter = 0;
        a.m = E->nrow;
        a.E = E;
        a.lambda = &lambda;
        a.Winit = W;
        const double* vs[1] = {V};
        const double* hs[1] = {Ht.data()};
        a.Vinit = vs;
        a.Hinit = hs;
        a.device = device;
        a.world = 1;
        lgr_ctx* raw = nullptr;
        create(&a, &raw);
        std::unique_ptr<lgr_ctx> c(raw);
        AllocStreamScope alloc_scope(c->st);
        *err = objective(c.get());
    });
}

}  // extern "C"

// ---- preprocessing scans (lgr_prep.cu) behind HOST buffers ------------------------------------------------------
namespace {
struct PrepStream {
    cudaStream_t st = nullptr;
    explicit PrepStream(int32_t device) {
        check_device();
        if (device >= 0) LGR_CUDA(cudaSetDevice(device));
        LGR_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    }
    ~PrepStream() {
        if (st) {
            cudaStreamSynchronize(st);
            cudaStreamDestroy(st);
        }
    }
};
size_t check_csc(int64_t nrow, int64_t ncol, const int32_t* colptr, const int32_t* rowidx, const double* values) {
    LGR_REQUIRE(ncol >= 0 && ncol < (int64_t)0x7fffffff && nrow >= 0 && nrow < (int64_t)0x7fffffff, "dimension out of range");
    LGR_REQUIRE(colptr, "null column pointers");
    LGR_REQUIRE(colptr[0] == 0, "colptr[0] must be 0");
    for (int64_t j = 0; j < ncol; ++j) LGR_REQUIRE(colptr[j + 1] >= colptr[j], "CSC column pointers must be non-decreasing");
    const size_t nnz = (size_t)colptr[ncol];
    LGR_REQUIRE(nnz == 0 || values, "null values");
    if (rowidx)
        for (size_t t = 0; t < nnz; ++t) LGR_REQUIRE(rowidx[t] >= 0 && rowidx[t] < nrow, "row index out of range");
    return nnz;
}
}  // namespace

extern "C" {

int lgr_normalize_csc(int64_t ncol, const int32_t* colptr, const double* values, double* out, double* colsums,
                      int32_t device) {
    return guarded([&] {
        const size_t nnz = check_csc(1, ncol, colptr, nullptr, values);
        LGR_REQUIRE(nnz == 0 || out, "null output");
        if (ncol == 0) return;
        PrepStream ps(device);
        AllocStreamScope alloc_scope(ps.st);
        DBuf<int> dp((size_t)ncol + 1);
        DBuf<double> dx(std::max<size_t>(nnz, 1)), dout(std::max<size_t>(nnz, 1)), dcs((size_t)ncol);
        LGR_CUDA(cudaMemcpyAsync(dp.p, colptr, ((size_t)ncol + 1) * sizeof(int), cudaMemcpyHostToDevice, ps.st));
        if (nnz) LGR_CUDA(cudaMemcpyAsync(dx.p, values, nnz * sizeof(double), cudaMemcpyHostToDevice, ps.st));
        prep_col_normalize((int)ncol, dp.p, dx.p, dout.p, dcs.p, ps.st);
        if (nnz) LGR_CUDA(cudaMemcpyAsync(out, dout.p, nnz * sizeof(double), cudaMemcpyDeviceToHost, ps.st));
        if (colsums) LGR_CUDA(cudaMemcpyAsync(colsums, dcs.p, (size_t)ncol * sizeof(double), cudaMemcpyDeviceToHost, ps.st));
        LGR_CUDA(cudaStreamSynchronize(ps.st));
    });
}

int lgr_row_vars_sparse(int64_t nrow, int64_t ncol_x, const int32_t* colptr, const int32_t* rowidx, const double* values,
                        const double* means_in, double ncol, double* means_out, double* vars_out, int32_t device) {
    return guarded([&] {
        LGR_REQUIRE(rowidx || (colptr && colptr[ncol_x >= 0 ? ncol_x : 0] == 0), "null row indices");
        const size_t nnz = check_csc(nrow, ncol_x, colptr, rowidx, values);
        if (nrow == 0) return;
        PrepStream ps(device);
        AllocStreamScope alloc_scope(ps.st);
        DBuf<int> di(std::max<size_t>(nnz, 1));
        DBuf<double> dx(std::max<size_t>(nnz, 1)), dm((size_t)nrow), dv((size_t)nrow), dc((size_t)nrow), dmi;
        if (nnz) {
            LGR_CUDA(cudaMemcpyAsync(di.p, rowidx, nnz * sizeof(int), cudaMemcpyHostToDevice, ps.st));
            LGR_CUDA(cudaMemcpyAsync(dx.p, values, nnz * sizeof(double), cudaMemcpyHostToDevice, ps.st));
        }
        if (means_in) {
            dmi.alloc((size_t)nrow);
            LGR_CUDA(cudaMemcpyAsync(dmi.p, means_in, (size_t)nrow * sizeof(double), cudaMemcpyHostToDevice, ps.st));
        }
        prep_row_mean_var((int)nrow, (int)ncol_x, nnz, di.p, dx.p, means_in ? dmi.p : nullptr, ncol, dm.p, dv.p, dc.p, ps.st);
        if (means_out) LGR_CUDA(cudaMemcpyAsync(means_out, dm.p, (size_t)nrow * sizeof(double), cudaMemcpyDeviceToHost, ps.st));
        if (vars_out) LGR_CUDA(cudaMemcpyAsync(vars_out, dv.p, (size_t)nrow * sizeof(double), cudaMemcpyDeviceToHost, ps.st));
        LGR_CUDA(cudaStreamSynchronize(ps.st));
    });
}

int lgr_scale_not_center_by_row(int64_t nrow, int64_t ncol, const int32_t* colptr, const int32_t* rowidx,
                                const double* values, double* out, double* rms_out, int32_t device) {
    return guarded([&] {
        LGR_REQUIRE(rowidx || (colptr && colptr[ncol >= 0 ? ncol : 0] == 0), "null row indices");
        const size_t nnz = check_csc(nrow, ncol, colptr, rowidx, values);
        LGR_REQUIRE(nnz == 0 || out, "null output");
        if (nrow == 0) return;
        PrepStream ps(device);
        AllocStreamScope alloc_scope(ps.st);
        DBuf<int> di(std::max<size_t>(nnz, 1));
        DBuf<double> dx(std::max<size_t>(nnz, 1)), dout(std::max<size_t>(nnz, 1)), drms((size_t)nrow);
        if (nnz) {
            LGR_CUDA(cudaMemcpyAsync(di.p, rowidx, nnz * sizeof(int), cudaMemcpyHostToDevice, ps.st));
            LGR_CUDA(cudaMemcpyAsync(dx.p, values, nnz * sizeof(double), cudaMemcpyHostToDevice, ps.st));
        }
        prep_scale_rows((int)nrow, (int)ncol, nnz, di.p, dx.p, dout.p, drms.p, ps.st);
        if (nnz) LGR_CUDA(cudaMemcpyAsync(out, dout.p, nnz * sizeof(double), cudaMemcpyDeviceToHost, ps.st));
        if (rms_out) LGR_CUDA(cudaMemcpyAsync(rms_out, drms.p, (size_t)nrow * sizeof(double), cudaMemcpyDeviceToHost, ps.st));
        LGR_CUDA(cudaStreamSynchronize(ps.st));
    });
}

}  // extern "C"

// ---- device-resident preprocessing pipeline -------------------------------------------------------------------------
struct PrepDataset {
    int nrow = 0, ncol = 0;
    size_t nnz = 0;
    DBuf<int> colptr, rowidx;
    DBuf<double> val;                 // raw counts, then the normalised values (in place)
    bool normalized = false;
    // scaleData (features x cells)
    int m = 0;
    size_t snnz = 0;
    DBuf<int> scolptr, srow;
    DBuf<double> sval64;
    DBuf<float> sval;
    bool scaled = false;
    // count structure of scaleData (values = count / library size / row rms): handed to lgr_inmf with the device view so
    // that the sweeps can run on the tensor cores (lgr_dense.cu); only when every raw value is an integer in 1..256
    bool counts_ok = false;
    DBuf<double> inv_lib, inv_rms;
};
struct lgr_prep {
    StreamHolder stream_holder;
    cudaStream_t st = nullptr;
    int device = 0;
    std::vector<PrepDataset> ds;
};

extern "C" {

int lgr_prep_create(int32_t d, const lgr_csc* raw, int32_t device, lgr_prep** out) {
    return guarded([&] {
        LGR_REQUIRE(out && raw && d >= 1, "null argument");
        *out = nullptr;
        check_device();
        if (device >= 0) LGR_CUDA(cudaSetDevice(device));
        std::unique_ptr<lgr_prep> p(new lgr_prep);
        LGR_CUDA(cudaGetDevice(&p->device));
        LGR_CUDA(cudaStreamCreateWithFlags(&p->st, cudaStreamNonBlocking));
        p->stream_holder.s = p->st;
        AllocStreamScope alloc_scope(p->st);
        p->ds.resize(d);
        for (int i = 0; i < d; ++i) {
            const lgr_csc& M = raw[i];
            LGR_REQUIRE(M.location == 0, "raw counts must be host matrices");
            // (row indices are range-checked on the device below: a host scan of 1e8 entries costs more than the upload)
            const size_t nnz = check_csc(M.nrow, M.ncol, M.colptr, nullptr, M.value_dtype == LGR_F64 ? static_cast<const double*>(M.values) : nullptr);
            LGR_REQUIRE(M.value_dtype == LGR_F64, "raw counts must be double (dgCMatrix @x)");
            LGR_REQUIRE(nnz == 0 || M.rowidx, "null row indices");
            PrepDataset& D = p->ds[i];
            D.nrow = (int)M.nrow;
            D.ncol = (int)M.ncol;
            D.nnz = nnz;
            D.colptr.alloc((size_t)D.ncol + 1);
            D.rowidx.alloc(std::max<size_t>(nnz, 1));
            D.val.alloc(std::max<size_t>(nnz, 1));
            LGR_CUDA(cudaMemcpyAsync(D.colptr.p, M.colptr, ((size_t)D.ncol + 1) * sizeof(int), cudaMemcpyHostToDevice, p->st));
            if (nnz) {
                LGR_CUDA(cudaMemcpyAsync(D.rowidx.p, M.rowidx, nnz * sizeof(int), cudaMemcpyHostToDevice, p->st));
                LGR_CUDA(cudaMemcpyAsync(D.val.p, M.values, nnz * sizeof(double), cudaMemcpyHostToDevice, p->st));
            }
        }
        DBuf<int> bad(1);
        bad.zero(p->st);
        for (auto& D : p->ds) prep_check_rows(D.nnz, D.rowidx.p, D.nrow, bad.p, p->st);
        int hbad = 0;
        LGR_CUDA(cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, p->st));
        LGR_CUDA(cudaStreamSynchronize(p->st));
        LGR_REQUIRE(hbad == 0, "row index out of range");
        for (auto& D : p->ds) {
            bad.zero(p->st);
            prep_check_counts(D.nnz, D.val.p, bad.p, p->st);
            LGR_CUDA(cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, p->st));
            LGR_CUDA(cudaStreamSynchronize(p->st));
            D.counts_ok = hbad == 0;
        }
        *out = p.release();
    });
}

int lgr_prep_normalize(lgr_prep* p) {
    return guarded([&] {
        LGR_REQUIRE(p, "null handle");
        LGR_CUDA(cudaSetDevice(p->device));
        for (auto& D : p->ds) {
            if (D.normalized) continue;
            AllocStreamScope alloc_scope(p->st);
            D.inv_lib.alloc((size_t)std::max(D.ncol, 1));
            prep_col_normalize(D.ncol, D.colptr.p, D.val.p, D.val.p, D.inv_lib.p, p->st);
            prep_reciprocal((size_t)D.ncol, D.inv_lib.p, D.inv_lib.p, p->st);
            D.normalized = true;
        }
        LGR_CUDA(cudaStreamSynchronize(p->st));
    });
}

int lgr_prep_gene_stats(lgr_prep* p, int32_t i, double* means, double* vars) {
    return guarded([&] {
        LGR_REQUIRE(p && i >= 0 && i < (int)p->ds.size(), "bad dataset index");
        LGR_CUDA(cudaSetDevice(p->device));
        AllocStreamScope alloc_scope(p->st);
        PrepDataset& D = p->ds[i];
        LGR_REQUIRE(D.normalized, "call lgr_prep_normalize first");
        if (D.nrow == 0) return;
        DBuf<double> dm((size_t)D.nrow), dv((size_t)D.nrow), dc((size_t)D.nrow);
        prep_row_mean_var(D.nrow, D.ncol, D.nnz, D.rowidx.p, D.val.p, nullptr, (double)D.ncol, dm.p, dv.p, dc.p, p->st);
        if (means) LGR_CUDA(cudaMemcpyAsync(means, dm.p, (size_t)D.nrow * sizeof(double), cudaMemcpyDeviceToHost, p->st));
        if (vars) LGR_CUDA(cudaMemcpyAsync(vars, dv.p, (size_t)D.nrow * sizeof(double), cudaMemcpyDeviceToHost, p->st));
        LGR_CUDA(cudaStreamSynchronize(p->st));
    });
}

int lgr_prep_scale(lgr_prep* p, int32_t i, int64_t m, const int32_t* feature_rows) {
    return guarded([&] {
        LGR_REQUIRE(p && i >= 0 && i < (int)p->ds.size(), "bad dataset index");
        LGR_REQUIRE(m >= 1 && m < (int64_t)0x7fffffff && feature_rows, "bad feature list");
        LGR_CUDA(cudaSetDevice(p->device));
        AllocStreamScope alloc_scope(p->st);
        PrepDataset& D = p->ds[i];
        LGR_REQUIRE(D.normalized, "call lgr_prep_normalize first");
        std::vector<int> lut((size_t)std::max(D.nrow, 1), -1);
        for (int64_t f = 0; f < m; ++f) {
            LGR_REQUIRE(feature_rows[f] >= 0 && feature_rows[f] < D.nrow, "feature row out of range");
            LGR_REQUIRE(lut[feature_rows[f]] < 0, "duplicate feature");
            lut[feature_rows[f]] = (int)f;
        }
        DBuf<int> dlut(lut.size()), cnt((size_t)D.ncol + 1);
        LGR_CUDA(cudaMemcpyAsync(dlut.p, lut.data(), lut.size() * sizeof(int), cudaMemcpyHostToDevice, p->st));
        cnt.zero(p->st);
        prep_subset_count(D.ncol, D.colptr.p, D.rowidx.p, dlut.p, cnt.p, p->st);
        view_forget(D.scolptr.p);   // views of the previous scaling die with its buffers
        D.scaled = false;
        D.scolptr.alloc((size_t)D.ncol + 1);
        dev_exclusive_scan_int(cnt.p, D.scolptr.p, (size_t)D.ncol + 1, p->st);
        int total = 0;
        LGR_CUDA(cudaMemcpyAsync(&total, D.scolptr.p + D.ncol, sizeof(int), cudaMemcpyDeviceToHost, p->st));
        LGR_CUDA(cudaStreamSynchronize(p->st));   // also keeps `lut` alive until its copy has run
        D.m = (int)m;
        D.snnz = (size_t)total;
        D.srow.alloc(std::max<size_t>(D.snnz, 1));
        D.sval64.alloc(std::max<size_t>(D.snnz, 1));
        D.sval.alloc(std::max<size_t>(D.snnz, 1));
        DBuf<double> sub(std::max<size_t>(D.snnz, 1)), rms((size_t)m);
        prep_subset_fill(D.ncol, D.colptr.p, D.rowidx.p, D.val.p, dlut.p, D.scolptr.p, D.srow.p, sub.p, p->st);
        prep_scale_rows((int)m, D.ncol, D.snnz, D.srow.p, sub.p, D.sval64.p, rms.p, p->st);
        prep_f64_to_f32(D.sval64.p, D.sval.p, D.snnz, p->st);
        D.inv_rms.alloc((size_t)m);
        prep_reciprocal((size_t)m, rms.p, D.inv_rms.p, p->st);
        LGR_CUDA(cudaStreamSynchronize(p->st));
        D.scaled = true;
    });
}

int lgr_prep_scaled_view(lgr_prep* p, int32_t i, lgr_csc* view) {
    return guarded([&] {
        LGR_REQUIRE(p && view && i >= 0 && i < (int)p->ds.size(), "bad dataset index");
        PrepDataset& D = p->ds[i];
        LGR_REQUIRE(D.scaled, "call lgr_prep_scale first");
        view->nrow = D.m;
        view->ncol = D.ncol;
        view->colptr = D.scolptr.p;
        view->rowidx = D.srow.p;
        view->values = D.sval.p;
        view->value_dtype = LGR_F32;
        view->location = LGR_CSC_DEVICE;
        view->row_scale = D.counts_ok ? D.inv_rms.p : nullptr;
        view->col_scale = D.counts_ok ? D.inv_lib.p : nullptr;
        view_register(view->colptr, view->nrow, view->ncol);
    });
}

int lgr_prep_scaled_download(lgr_prep* p, int32_t i, int32_t* colptr, int32_t* rowidx, double* values) {
    return guarded([&] {
        LGR_REQUIRE(p && i >= 0 && i < (int)p->ds.size(), "bad dataset index");
        LGR_CUDA(cudaSetDevice(p->device));
        PrepDataset& D = p->ds[i];
        LGR_REQUIRE(D.scaled, "call lgr_prep_scale first");
        if (colptr) LGR_CUDA(cudaMemcpyAsync(colptr, D.scolptr.p, ((size_t)D.ncol + 1) * sizeof(int), cudaMemcpyDeviceToHost, p->st));
        if (rowidx && D.snnz) LGR_CUDA(cudaMemcpyAsync(rowidx, D.srow.p, D.snnz * sizeof(int), cudaMemcpyDeviceToHost, p->st));
        if (values && D.snnz) LGR_CUDA(cudaMemcpyAsync(values, D.sval64.p, D.snnz * sizeof(double), cudaMemcpyDeviceToHost, p->st));
        LGR_CUDA(cudaStreamSynchronize(p->st));
    });
}

void lgr_prep_destroy(lgr_prep* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    {
        AllocStreamScope alloc_scope(p->st);
        for (auto& D : p->ds) view_forget(D.scolptr.p);
        p->ds.clear();
    }
    delete p;
}

}  // extern "C"
