// liger_b200 -- shared declarations of the CUDA implementation (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <stdexcept>
#include <cmath>
#include <string>
#include <vector>

#include "../../include/liger_b200.h"

namespace lgr {

// ---------------------------------------------------------------------------------------------
// error handling: C++ exceptions inside the library, converted to status codes at the C ABI.
// ---------------------------------------------------------------------------------------------
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string& msg);

#define LGR_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            throw ::lgr::Error(LGR_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) +  \
                                                 " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
    } while (0)

#define LGR_REQUIRE(cond, msg)                                          \
    do {                                                                \
        if (!(cond)) throw ::lgr::Error(LGR_ERR_INVALID, std::string(msg)); \
    } while (0)

// Stream the RAII buffers below allocate / free on (stream-ordered allocation from the device's default memory
// pool: no device-wide synchronisation, and the pool keeps freed blocks for the next call).
cudaStream_t& alloc_stream();
struct AllocStreamScope {
    cudaStream_t prev;
    explicit AllocStreamScope(cudaStream_t s) : prev(alloc_stream()) { alloc_stream() = s; }
    ~AllocStreamScope() { alloc_stream() = prev; }
};

// RAII device buffer
template <typename T>
struct DBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaStream_t st = nullptr;
    DBuf() {}
    explicit DBuf(size_t n_) { alloc(n_); }
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    DBuf(DBuf&& o) noexcept : p(o.p), n(o.n), st(o.st) { o.p = nullptr; o.n = 0; }
    DBuf& operator=(DBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; st = o.st; o.p = nullptr; o.n = 0; }
        return *this;
    }
    void alloc(size_t n_) {
        release();
        n = n_;
        if (n) {
            st = alloc_stream();
            cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&p), n * sizeof(T), st);
            if (e != cudaSuccess) {
                p = nullptr; n = 0;
                cudaGetLastError();
                throw Error(LGR_ERR_ALLOC, std::string("cudaMallocAsync of ") + std::to_string(n_ * sizeof(T)) + " bytes: " + cudaGetErrorString(e));
            }
        }
    }
    void zero(cudaStream_t s) { if (n) LGR_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
    void release() { if (p) cudaFreeAsync(p, st); p = nullptr; n = 0; }
    ~DBuf() { release(); }
};

// ---------------------------------------------------------------------------------------------
// device-side description of one dataset (one rank's cell shard of it)
// ---------------------------------------------------------------------------------------------
// Layout in HBM (all row strides are KP = 32 or 64 floats so that one factor row is one or two
// 128-byte lines):
//   X (stacked [E_i; P_i], rows = m + u_i) is held twice:
//     * CSC   colptr[n+1], rowidx[nnz] (int32), val[nnz] (fp32)        -> K1  B = L^T X   (gather of L rows)
//     * tiled CSR: cells are cut into tiles of TC; inside a tile entries are ordered by gene;
//       tptr[ntiles*rows + 1], tcell[nnz] (uint16, cell index inside the tile), tval[nnz] (fp32)
//                                                                   -> K2  R = X H     (gather of H rows)
//   L = [W + V_i; U_i]  rows x KP fp32        H, B  n x KP fp32       masks n x uint64
//   V64 rows x KP fp64 (V_i stacked over U_i) R rows x KP fp32        G KP x KP fp64
struct DsDev {
    int n;       // local cells
    int rows;    // m + u
    int m;       // shared rows
    int u;       // unshared rows
    int ntiles;  // cell tiles of TC
    int ltx_blk_off;  // first flat block of this dataset in k_spmm_ltx
    int xh_blk_off;   // first flat block (tile) in k_spmm_xh
    int prep_blk_off; // first flat block in k_prep
    const int* colptr;
    const int* rowidx;
    const float* val;
    const int* cptr8;            // gene-tiled CSC (padded to 8): [ngt * n + 1] segment starts in groups of 8
    const unsigned short* crow;  // (row inside the gene tile) * KP/4
    const unsigned int* ctag;    // per group of 8: cell % CB
    const float* cval;
    const int* tptr8;            // cell-tiled CSR (padded to 8): [ntiles * rows + 1]
    const unsigned short* tcell; // (cell inside the cell tile) * KP/4
    const unsigned int* ttag;    // per group of 8: gene
    const float* tval;
    int ngt;                     // gene tiles
    int asplit_half;             // dense path, KP = 64: 16-bit words between the slab images of factors 0..31 and 32..63
    float* L;
    float* H;
    float* B;
    float* U;            // u = CtC^-1 b of the H solve, written by the fused epilogue of k_spmm_ltx_tiled (else null)
    const double* Pinv;  // CtC^-1 of the H solve (KP x KP fp64, k_gram_inv)
    unsigned long long* hmask;
    double* V;      // rows x KP
    double* CtBv;   // rows x KP   rhs of the V/U solve
    unsigned long long* vmask;
    float* R;       // rows x KP   (inside the all-reduce arena)
    double* G;      // KP x KP     H^T H (inside the all-reduce arena)
    double* LtL;    // KP x KP     L~^T L~
    double* VtV;    // KP x KP     V~^T V~
    double lambda;
    double sqnorm;  // ||[E;P]||_F^2 (global, after all-reduce)
    // dense-tile path (lgr_dense.cu): k_prep also writes the bf16 3-term split of rs[g] * L~[g, :] in the operand-slab
    // image k_dense_ltx copies into shared memory (null in the gather path)
    unsigned short* Asplit;
    const float* rs;   // rows: row scale of the count factorisation X = N * rs * cs
};

// device-side description of one dataset for the dense-tile sweeps (lgr_dense.cu)
struct DenseDs {
    int n, rows;
    int ntile1;       // 256-cell tiles (k_dense_ltx)
    int nstage1;      // 64-gene stages
    int k1_tile_off;  // first flat tile of this dataset
    int ngb;          // 512-gene blocks (k_dense_xh)
    int ncstage;      // 32-cell stages
    int gh_blk_off;   // first flat block in k_gram_h
    int ldf;          // row stride (floats) of B, H and R, (doubles) of G: KP.  For KP = 64 every dataset has two descriptors,
    int kloc;         // one per half of the factors (32 operand rows x 3 terms each); kloc = factors of this half
    long long k2_flat_off;   // first flat (gene block, cell stage) index of this dataset
    const int* seg1;
    const unsigned int* ent1;
    const unsigned short* Asplit;
    const int* seg2;
    const unsigned int* ent2;
    const float* rs;   // rows
    const float* cs;   // n
    float* B;
    const float* H;
    float* R;
    double* G;
    unsigned short* Hs;   // bf16 split image of cs * H (A slabs of k_dense_xh), written by k_gram_tc: 6 KB per 32 cells
};

// One batch of NNLS columns sharing a Gram matrix (gram / ginv / ok are produced by k_gram_inv)
struct NnlsGroup {
    const double* gram;         // KP x KP  CtC
    const double* ginv;         // KP x KP  CtC^-1 (valid when *ok != 0)
    const int* ok;
    const void* rhs;            // T: ncols x ld  (column j at rhs + j*ld)
    const void* u;              // optional T: ncols x ld, u_j = CtC^-1 rhs_j precomputed on the tensor cores (lgr_tc.cu)
    void* x;                    // T: ncols x ld
    unsigned long long* mask;   // ncols, warm start in/out (never null)
    int ncols;
    int ld;
    long long col_off;          // first column of this group in the launch's flat column range
};

// CtC = cA * A + cB * B (A, B fp64 KP x KP, B may be null) -> gram, ginv, ok
struct GramSpec {
    const double* A;
    const double* B;
    double cA, cB;
    double* gram;
    double* ginv;
    int* ok;
};

struct Counters {
    unsigned long long nnls_pivots;
    unsigned long long nnls_unconverged;
    unsigned long long nnls_bad_pivot;
    unsigned long long nnls_spill;  // solves that used the global-memory factor scratch
    unsigned long long hist[40];    // histogram of pivots per column (last bin: >= 39)
};

// ---------------------------------------------------------------------------------------------
// kernel launch wrappers (lgr_kernels.cu)
// ---------------------------------------------------------------------------------------------
inline int prep_rows_per_block(int kp) { return 2048 / kp; }  // keeps k_prep's static smem < 48 KB

struct NnlsConfig {
    int threads;      // threads per CTA (warps x 32; every warp owns 32 columns at a time)
    int pool_elems;   // per-warp shared-memory pool for the per-column Cholesky factors
    size_t smem;
    int grid_cap;     // persistent grid size
    int ct_cols;      // (unused)
    int ctas_per_sm;
};
NnlsConfig nnls_config(int k, int kp, bool is_double, int num_sms);

void launch_prep(const DsDev* ds, int nds, int total_blocks, const double* W, int k, int kp, cudaStream_t st);
void launch_spmm_ltx_tiled(const DsDev* ds, int nds, int total_blocks, int kp, int cb, int gt, bool fuse_u, cudaStream_t st);
size_t spmm_ltx_smem(int kp, int cb, int gt, bool fuse_u);
void spmm_init();
void launch_spmm_xh(const DsDev* ds, int nds, int total_tiles, int k, int kp, int tc, cudaStream_t st);
void launch_gram_inv(const GramSpec* specs, int nspecs, int k, int kp, cudaStream_t st);
void launch_nnls(bool is_double, const NnlsGroup* groups, int ngroups, long long total_cols, int k, int kp,
                 const NnlsConfig& cfg, Counters* cnt, int flags, cudaStream_t st);
void nnls_init();
// register-resident fp32 path (lgr_nnls_reg.cu): k <= 62.  Groups with a singular Gram are skipped and columns whose
// system exceeds 16 unknowns are flagged in their warm-start mask: run launch_nnls with flags 16 | 64 right after it
bool nnls_reg_supported(int k, int kp);
void launch_nnls_reg(const NnlsGroup* groups, int ngroups, long long total_cols, int k, int kp, int num_sms, Counters* cnt,
                     int flags, cudaStream_t st);
void nnls_reg_init();
// continues the columns k_nnls_reg<64> flagged as too big for its register templates (lgr_nnls_wbig.cu)
void launch_nnls_wbig(const NnlsGroup* groups, int ngroups, long long total_cols, int k, int num_sms, Counters* cnt,
                      cudaStream_t st);
void nnls_wbig_init();
// the same hand-over, thread per column with the principal block in shared memory (lgr_nnls_scol.cu): the default
void launch_nnls_scol(const NnlsGroup* groups, int ngroups, long long total_cols, int k, int num_sms, Counters* cnt, cudaStream_t st);
void nnls_scol_init();
// systems of 17..26 unknowns: two lanes per column, the principal block in their registers (lgr_nnls_pair.cu)
void launch_nnls_pair(const NnlsGroup* groups, int ngroups, long long total_cols, int k, int num_sms, Counters* cnt, cudaStream_t st);
void nnls_pair_init();
bool nnls_wbig64_supported(int k, int kp, long long total_cols, int num_sms);
void launch_nnls_wbig64(const NnlsGroup* groups, int ngroups, long long total_cols, int k, int num_sms, Counters* cnt, int flags,
                        cudaStream_t st);
// warp-per-column FP64 path for small batches (lgr_nnls_warp.cu): k <= 32; singular groups are skipped (flag 16 mop-up)
bool nnls_warp_supported(int k, int kp, long long total_cols, int num_sms);
void launch_nnls_warp(const NnlsGroup* groups, int ngroups, long long total_cols, int k, int num_sms, Counters* cnt, int flags,
                      cudaStream_t st);
void launch_pb_tc(const NnlsGroup* groups, int ngroups, long long total_tiles, int kp, int num_sms, cudaStream_t st);
void tc_init();
void launch_vrhs(const DsDev* ds, int nds, const double* W, int k, int kp, int max_rows, cudaStream_t st);
void launch_wrhs(const DsDev* ds, int nds, int m, double* CtBw, double* Gsum, int k, int kp, cudaStream_t st);
void launch_pack_gram(const double* G, float* tail, int n, cudaStream_t st);
void launch_unpack_gram(double* G, const float* tail, int n, cudaStream_t st);
void launch_objective(const DsDev* ds, int nds, int k, int kp, double* obj_slot, cudaStream_t st);
size_t spmm_xh_smem(int kp, int tc);
void kernels_init();  // opt-in shared memory attributes of every kernel, once per device (safe to call repeatedly)

// dense-tile tensor-core sweeps (lgr_dense.cu)
struct DenseLayout {
    int ntile1 = 0, nstage1 = 0, ngb = 0, ncstage = 0;
    DBuf<int> seg1, seg2;
    DBuf<unsigned int> ent1, ent2;
    DBuf<unsigned short> Asplit, Hs;
    DBuf<float> rs, cs;
};
void dense_build(int n, int rows, const int* colptr, const int* rowidx, const float* val, size_t nnz, const float* rs, const float* cs,
                 int halves, DenseLayout& out, int* bad, cudaStream_t st);
void launch_gram_cross(const DenseDs* ds, int nds, int max_n, cudaStream_t st);
void launch_dense_ltx(const DenseDs* ds, int nds, int total_tiles, int ldf, int num_sms, cudaStream_t st);
void launch_dense_xh(const DenseDs* ds, int nds, long long total_flat, int ldf, int num_sms, cudaStream_t st);
void launch_gram_h(const DenseDs* ds, int nds, int total_items, int ldf, int num_sms, cudaStream_t st);
int dense_gram_tile();
void dense_init();

// copies between caller host memory and the device (lgr_stage.cu): pageable buffers go through page-locked bounce buffers
bool host_is_pinned(const void* p);
void copy_in(void* dst, const void* src, size_t bytes, cudaStream_t cs, bool pinned_promise);
void copy_out(void* dst, const void* src, size_t bytes, cudaStream_t st, bool pinned_promise);

// layout / conversion kernels (lgr_layout.cu)
void dev_f64_to_f32(const double* in, float* out, size_t n, cudaStream_t st);
void dev_add_int(int* p, size_t n, int delta, cudaStream_t st);
void dev_sumsq(const float* v, size_t n, double* out, cudaStream_t st);
void dev_stack_csc(int n, const int* pE, const int* iE, const float* vE, const int* pP, const int* iP,
                   const float* vP, int m, int* pS, int* iS, float* vS, cudaStream_t st);
void dev_build_padded(int mode, int n, int rows, int tile, int scale, int tagmod, const int* colptr, const int* rowidx,
                      const float* val, size_t nnz, DBuf<int>& ptr8, DBuf<unsigned short>& oloc, DBuf<unsigned int>& otag,
                      DBuf<float>& oval,
                      cudaStream_t st);
void dev_colmajor_to_padded(const double* in, double* out, int rows, int k, int kp, cudaStream_t st);   // R m x k -> [row][KP]
void dev_padded_to_colmajor(const double* in, double* out, int rows, int k, int kp, cudaStream_t st);
void dev_hpad_to_colmajor(const float* H, double* out, int n, int k, int kp, cudaStream_t st);          // [cell][KP] f32 -> n x k f64
void dev_colmajor_to_hpad(const double* in, float* H, int n, int k, int kp, cudaStream_t st);
void dev_mask_from_positive(const float* H, unsigned long long* mask, int n, int k, int kp, cudaStream_t st);


// preprocessing scans (lgr_prep.cu)
void prep_col_normalize(int ncol, const int* colptr, const double* x, double* out, double* colsum, cudaStream_t st);
void prep_row_mean_var(int nrow, int ncol_x, size_t nnz, const int* rowidx, const double* x, const double* means_in,
                       double ncol, double* means_out, double* vars_out, double* scratch_count, cudaStream_t st);
void prep_scale_rows(int nrow, int ncol, size_t nnz, const int* rowidx, const double* x, double* out, double* rms,
                     cudaStream_t st);
void prep_subset_count(int ncol, const int* colptr, const int* rowidx, const int* lut, int* count, cudaStream_t st);
void prep_subset_fill(int ncol, const int* colptr, const int* rowidx, const double* x, const int* lut, const int* newptr,
                      int* newrow, double* newx, cudaStream_t st);
void prep_f64_to_f32(const double* in, float* out, size_t n, cudaStream_t st);
void prep_check_rows(size_t nnz, const int* rowidx, int nrow, int* bad, cudaStream_t st);
void prep_check_counts(size_t nnz, const double* x, int* bad, cudaStream_t st);
void prep_reciprocal(size_t n, const double* in, double* out, cudaStream_t st);
void prep_check_same(size_t n, const double* a, const double* b, int* bad, cudaStream_t st);   // bad |= 1 when a != b anywhere
void dev_exclusive_scan_int(const int* in, int* out, size_t n, cudaStream_t st);   // lgr_layout.cu (cub)

// lgr_online.cu: kernels of online iNMF (minibatch H solves, running statistics, HALS sweeps)
struct OnlineDs {
    const float* H;      // minibatch H rows (nb x 32)
    int nb;              // minibatch size b_i
    double* A;           // 32 x 32
};
struct HalsDs {
    double* V;           // m x 32
    const double* A;     // 32 x 32
    const double* B;     // m x 32
};
struct OnlineX {          // one dataset's resident fp32 CSC and its per-minibatch constants
    const int* colptr;
    const int* rowidx;
    const float* val;
    const float* Lt;      // L~_i, m x 32
    double* Bstat;        // B_i, m x 32
    double inv_b;         // 1 / b_i
    int off;              // first minibatch cell of this dataset in the flat [0, mb_total) range
    int pad0;
};
void launch_online_ltx_all(const OnlineX* xs, int d, const int* sched, int mb_total, const int* t_dev, float* B, cudaStream_t st);
void launch_online_stats_all(const OnlineDs* ds, const OnlineX* xs, int d, int m, int k, const int* sched, int mb_total, int* t_dev,
                             const float* H, cudaStream_t st);
void launch_online_next(int* t_dev, cudaStream_t st);
void launch_online_gram(const double* W, const double* const* V, int d, int m, int k, float* const* Lt, double* const* LtL,
                        double* const* VtV, cudaStream_t st);
void launch_online_ltx(const int* colptr, const int* rowidx, const float* val, const int* cells, int nb, const float* Lt, float* B,
                       cudaStream_t st);
void launch_online_stats_a(const OnlineDs* ds, int d, int k, double s, cudaStream_t st);
void launch_online_hals(double* W, const HalsDs* ds, int d, int m, int k, double lambda, int sweeps, int first_v, cudaStream_t st);
void launch_online_cross(const float* H, const float* B, size_t n, double* out, cudaStream_t st);
void launch_online_sumsq(const float* x, size_t n, double* out, cudaStream_t st);
// lgr_centroid.cu: centroidAlign on device-resident FP64 H (column-major n_i x k, read only) -> out (n_i x kd)
void centroid_align_core(int d, int k, const int* n, const double* const* Hdev, const lgr_ca_params& p, double* const* out_dev,
                         int* const* raw_max, int* const* z_max, int* const* r_max, cudaStream_t st);
// lgr_align.cu: quantileNorm on device-resident FP64 H (column-major n_i x k, replaced by H.norm)
void quantile_norm_core(int d, int k, const int* n, double* const* Hdev, const lgr_qn_params& p, int* const* labels_dev,
                        lgr_qn_info* info, cudaStream_t st);

// R >= 3.6 sample(): R_unif_index by rejection over ceil(log2(n)) bits assembled from 16-bit chunks of unif_rand()
struct RSampler {
    lgr_rng* rng;
    double unif() {
        double u;
        lgr_r_runif(rng, 1, 0.0, 1.0, &u);
        return u;
    }
    long long unif_index(long long dn) {
        if (dn <= 0) return 0;
        const int bits = (int)std::ceil(std::log2((double)dn));
        for (;;) {
            long long v = 0;
            for (int n = 0; n <= bits; n += 16) {
                const int v1 = (int)std::floor(unif() * 65536);
                v = 65536 * v + v1;
            }
            if (bits < 64) v &= ((1LL << bits) - 1);
            if (v < dn) return v;
        }
    }
    // positions of sample.int(n, size) without replacement; `out` may be null (the stream is advanced all the same)
    void sample(int n, int size, std::vector<int>& scratch, int* out) {
        if (size < 2) {
            for (int i = 0; i < size; ++i) {
                const int j = (int)unif_index(n);
                if (out) out[i] = j;
            }
            return;
        }
        scratch.resize(n);
        for (int i = 0; i < n; ++i) scratch[i] = i;
        int nn = n;
        for (int i = 0; i < size; ++i) {
            const int j = (int)unif_index(nn);
            if (out) out[i] = scratch[j];
            scratch[j] = scratch[--nn];
        }
    }
};


}  // namespace lgr
