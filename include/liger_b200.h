/*
 * liger_b200 -- C ABI of the B200-native iNMF / UINMF alternating-NNLS path.
 *
 * This header is the drop-in boundary.  Every entry point
 * replaces one native call the reference makes on this path (citations are
 * into the reference tree, welch-lab/liger):
 *
 *   lgr_inmf            <- RcppPlanc::inmf(objectList, k, lambda, niter, Hinit, Vinit, Winit, nCores, verbose)
 *                          call sites R/integration.R:438-441, R/cINMF.R:294-297, R/suggestK.R:78-88
 *   lgr_uinmf           <- RcppPlanc::uinmf(object, unsharedList, k, lambda, niter, nCores, verbose)
 *                          call site  R/integration.R:1285-1287
 *   lgr_bppnnls_prod    <- RcppPlanc::bppnnls_prod(CtC, CtB, nCores)
 *                          call sites R/cINMF.R:481,489,501; R/optimizeNewParam.R:104,118,134,291,329
 *   lgr_objerr_i        <- objErr_i(H, W, V, E, lambda)      src/objErr.cpp:13-31 (glue src/RcppExports.cpp:270-284)
 *   lgr_r_set_seed / lgr_r_runif
 *                       <- set.seed(seed); runif(n, 0, 2)    R/integration.R:412-437 (seeded initialisation)
 *
 * Conventions
 *   - plain C, POD structs, no exceptions cross this boundary; every function that can fail
 *     returns 0 on success and a non-zero lgr_status otherwise, with a message in lgr_last_error().
 *   - all matrices are R layout: dense = column-major double; sparse = dgCMatrix slots
 *     (@p int32 column pointers, @i int32 0-based row indices, @x values, @Dim).
 *   - the library never retains or frees caller pointers; outputs are written only into
 *     caller-allocated buffers.  All device memory is owned by the library.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     LGR_ERR_NO_DEVICE.
 */
#ifndef LIGER_B200_H
#define LIGER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define LGR_API __attribute__((visibility("default")))
#else
#define LGR_API
#endif

#define LGR_VERSION_MAJOR 0
#define LGR_VERSION_MINOR 1
#define LGR_MAX_K 64 /* factors; passive sets are 64-bit masks */

typedef enum lgr_status {
    LGR_OK = 0,
    LGR_ERR_INVALID = 1,   /* bad argument (message says which) */
    LGR_ERR_NO_DEVICE = 2, /* no usable CUDA device: there is no CPU path */
    LGR_ERR_CUDA = 3,      /* CUDA runtime / kernel failure */
    LGR_ERR_NCCL = 4,      /* NCCL failure or libnccl not loadable */
    LGR_ERR_ALLOC = 5,
    LGR_ERR_INTERRUPT = 6  /* the caller's interrupt callback asked to stop */
} lgr_status;

typedef enum lgr_dtype { LGR_F64 = 0, LGR_F32 = 1 } lgr_dtype;

/* One dgCMatrix (genes x cells, CSC).  `values` is double (R) or float (pre-converted). */
typedef struct lgr_csc {
    int64_t nrow;
    int64_t ncol;
    const int32_t* colptr; /* ncol + 1 */
    const int32_t* rowidx; /* nnz, 0-based; need not be sorted within a column */
    const void* values;    /* nnz */
    int32_t value_dtype;   /* lgr_dtype */
    int32_t location;      /* 0: host pointers; LGR_CSC_DEVICE: device pointers (views returned by lgr_prep_scaled_view) */
    /* Optional count structure (both NULL = none).  The caller asserts that every stored value is
           values[t] = count * row_scale[row] * col_scale[col]      with an integer count in 1..256,
       which is what rliger's scaleData is: normalize() divides a count by its cell's library size (col_scale = 1 / colSums,
       R/preprocess.R:604) and scaleNotCenter() divides by the gene's root mean square (row_scale = 1 / rms,
       src/data_processing.cpp:42-58).  The library verifies the assertion entry by entry (LGR_ERR_INVALID if it does not
       hold) and then runs the two sweeps over X as dense-tile tensor-core contractions (csrc/lgr_dense.cu) instead of the
       shared-memory gathers.  Same results to fp32 rounding.  Used for k <= 32 (UINMF: when the unshared blocks carry their
       structure too, with the same col_scale as their shared block); ignored otherwise.
       row_scale: nrow doubles, col_scale: ncol doubles, host or device like the other arrays. */
    const double* row_scale;
    const double* col_scale;
} lgr_csc;
#define LGR_CSC_DEVICE 1

/* flags for lgr_inmf_args.flags */
#define LGR_FLAG_TRAJECTORY 0x1u  /* evaluate the objective after every iteration into out->obj_traj */
#define LGR_FLAG_COLD_START 0x2u  /* disable NNLS warm starts (every solve starts from the empty passive set) */
#define LGR_FLAG_PINNED_IO 0x4u   /* caller promises that every host buffer (inputs and outputs) is page-locked: the
                                     library copies straight from / into them.  Without the flag each buffer is probed
                                     (cudaPointerGetAttributes) and pageable ones -- R's heap, numpy arrays -- go through
                                     a ring of page-locked bounce buffers filled by a few host threads, so that the DMA
                                     engine still runs at PCIe speed */
#define LGR_FLAG_TENSOR_U 0x10u   /* form u = CtC^-1 b of the H solve on the tensor cores (tcgen05, 3xTF32 split, TMEM
                                     accumulators): as an epilogue of the B = L^T X sweep while the B tile is still in shared
                                     memory, or as a separate pass when the tile geometry does not allow it.  Same results
                                     to ~1e-5; at k = 30 it trades +0.11 ms in the sweep for -0.06 ms in the solver, so it is
                                     off by default */
#define LGR_FLAG_LEGACY_NNLS 0x20u /* verification only: run the H solves with the shared-memory-factor NNLS kernel
                                     (lgr_nnls.cu) instead of the register-resident one (lgr_nnls_reg.cu, k <= 32) */
#define LGR_FLAG_NO_DENSE 0x40u    /* ignore row_scale / col_scale: always run the gather sweeps */
#define LGR_FLAG_LOCAL_SHARD 0x8u /* world > 1: E/P already hold only THIS rank's cells (one process per GPU, each
                                     with its own shard); without it every rank passes the full matrices and the
                                     library takes columns [n*rank/world, n*(rank+1)/world) */

#define LGR_FLAG_DATASET_SHARD 0x80u /* world > 1: WHOLE-DATASET placement (the degenerate case of cell sharding when the
                                     datasets divide over the ranks): E/P/lambda/Vinit describe only THIS rank's datasets, with all
                                     their cells; V_i / U_i / H_i are solved and returned locally, and the only collective per
                                     iteration is the all-reduce of the W solve's normal equations (sum_i (R_i - G_i V_i^T) and
                                     sum_i G_i: (m + k) x k doubles) instead of every R_i and G_i.  W and objErr come back identical
                                     on every rank.  Explicit Winit / Vinit required (the seeded draw is defined on the global
                                     dataset order). */

/* Arguments of one iNMF / UINMF factorisation.  iNMF: P == NULL, Uinit == NULL. */
typedef struct lgr_inmf_args {
    int32_t d;     /* number of datasets */
    int32_t k;     /* factors, 1..LGR_MAX_K */
    int32_t niter; /* ANLS iterations (H, V, W per iteration) */
    uint32_t flags;
    int64_t m;             /* shared features (rows of every E[i]) */
    const lgr_csc* E;      /* d shared blocks, m x n_i */
    const lgr_csc* P;      /* NULL, or d unshared blocks u_i x n_i (nrow == 0 => none for that dataset) */
    const double* lambda;  /* d penalties (iNMF: the scalar repeated) */
    const double* Winit;   /* m x k; NULL (together with Vinit / Uinit) = draw the initial factors inside the library,
                              as RcppPlanc does when rliger passes Hinit = Vinit = Winit = NULL (R/suggestK.R:78-88):
                              R's Mersenne-Twister stream seeded with init_seed, U(0, 2), column-major, in the order of
                              .runINMF.list (R/integration.R:412-437): W, V_1..V_d, (U_1..U_d,) H_1..H_d */
    const double* const* Vinit; /* d pointers, each m x k (NULL: drawn, see Winit) */
    const double* const* Uinit; /* NULL, or d pointers, each u_i x k (NULL where u_i == 0) */
    const double* const* Hinit; /* NULL, or d pointers, each n_i x k.  H is solved first, so Hinit only
                                   matters when niter == 0 (as in RcppPlanc). */
    /* execution */
    int32_t device; /* CUDA ordinal; -1 = current device */
    int32_t rank;   /* cell-shard rank of this process, 0 when world == 1 */
    int32_t world;  /* number of processes sharing the factorisation (one GPU each) */
    int32_t init_seed; /* seed of the in-library initialisation when Winit / Vinit are NULL (0 is read as 1) */
    const void* nccl_unique_id; /* 128 bytes from lgr_nccl_unique_id() (same on all ranks); NULL if world == 1.
                                   Passing the SAME id again reuses the cached communicator (no ncclCommInitRank). */
    /* optional: polled between iterations; non-zero return aborts with LGR_ERR_INTERRUPT */
    int (*interrupt)(void* user);
    void* interrupt_user;
} lgr_inmf_args;

typedef struct lgr_timings {
    double upload_ms;   /* H2D + on-device layout conversion */
    double loop_ms;     /* the ANLS loop (device time, CUDA events) */
    double download_ms; /* D2H of factors */
    double h_phase_ms, vw_phase_ms, comm_ms;
    int64_t kernel_launches; /* kernels launched inside the loop */
    int64_t nnls_unconverged; /* columns that hit the pivot cap (expected 0) */
    int64_t nnls_pivots;      /* total BPP pivots over all H solves */
} lgr_timings;

/* Caller-allocated outputs.  Any pointer may be NULL to skip that output. */
typedef struct lgr_inmf_out {
    double* W;          /* m x k */
    double* const* V;   /* d pointers, m x k */
    double* const* U;   /* d pointers, u_i x k (UINMF) */
    double* const* H;   /* d pointers, n_i x k  (cell x factor, as RcppPlanc returns it; the R wrapper
                           transposes, R/integration.R:453) */
    uint64_t* const* H_passive; /* optional: d pointers, n_i passive-set bitmasks (bit f set <=> H[j,f] free) */
    double* obj_traj;   /* niter values when LGR_FLAG_TRAJECTORY */
    double objErr;      /* final objective (sum over datasets of objErr_i) */
    lgr_timings timings;
} lgr_inmf_out;

/* ---- one-shot calls with HOST buffers (what the Rcpp shim binds) ------------------------- */
LGR_API int lgr_inmf(const lgr_inmf_args* args, lgr_inmf_out* out);
LGR_API int lgr_uinmf(const lgr_inmf_args* args, lgr_inmf_out* out);

/* X = argmin_{X>=0} ||C X - B|| from CtC (k x k) and CtB (k x c); X is k x c.  FP64 on the device. */
LGR_API int lgr_bppnnls_prod(int32_t k, int64_t c, const double* CtC, const double* CtB, double* X, int32_t device);

/* objErr_i: H is k x n (as in src/objErr.cpp), W, V are m x k, E is m x n. */
LGR_API int lgr_objerr_i(int32_t k, const double* H, const double* W, const double* V, const lgr_csc* E, double lambda,
                 int32_t device, double* err);

/* ---- preprocessing that produces the path's input; FP64 on the device, HOST buffers --------
   The matrix is a dgCMatrix: colptr (ncol + 1), rowidx (nnz), values (nnz doubles); outputs keep the pattern. */
/* normalize.dgCMatrix (R/preprocess.R:592-608): out = x / colSums(x)[col]; colsums (ncol) may be NULL. */
LGR_API int lgr_normalize_csc(int64_t ncol, const int32_t* colptr, const double* values, double* out, double* colsums,
                              int32_t device);
/* rowVars_sparse_rcpp(x, means, ncol) (src/data_processing.cpp:152-172) together with the Matrix::rowMeans that
   feeds it (R/preprocess.R:1024): means_in may be NULL (then means = rowSums / ncol(x)); `ncol` is the divisor's
   column count (the full matrix when x is a column chunk).  means_out / vars_out: nrow doubles, either may be NULL. */
LGR_API int lgr_row_vars_sparse(int64_t nrow, int64_t ncol_x, const int32_t* colptr, const int32_t* rowidx,
                                const double* values, const double* means_in, double ncol, double* means_out,
                                double* vars_out, int32_t device);
/* scaleNotCenter_byRow_rcpp (src/data_processing.cpp:42-58): out = x / sqrt(rowSums(x^2) / (ncol - 1)), 0 for rows
   whose root mean square is 0; rms_out (nrow) may be NULL. */
LGR_API int lgr_scale_not_center_by_row(int64_t nrow, int64_t ncol, const int32_t* colptr, const int32_t* rowidx,
                                        const double* values, double* out, double* rms_out, int32_t device);

/* ---- device-resident preprocessing pipeline (second half) ----------------------------------
   raw counts -> normalize -> per-gene statistics (to the host, which selects the variable genes) -> scaleNotCenter ->
   scaleData that STAYS in HBM and feeds lgr_inmf / lgr_ctx_create through a device view: X is never copied back to the
   host nor uploaded a second time. */
typedef struct lgr_prep lgr_prep;
LGR_API int lgr_prep_create(int32_t d, const lgr_csc* raw /* d matrices, genes_i x n_i, host */, int32_t device, lgr_prep** out);
LGR_API int lgr_prep_normalize(lgr_prep* p);                      /* normalize.dgCMatrix, every dataset, in place */
LGR_API int lgr_prep_gene_stats(lgr_prep* p, int32_t i, double* means /* genes_i */, double* vars /* genes_i */);
/* scaleNotCenter of dataset i: rows feature_rows[0..m) (indices into its gene list, in feature order) */
LGR_API int lgr_prep_scale(lgr_prep* p, int32_t i, int64_t m, const int32_t* feature_rows);
LGR_API int lgr_prep_scaled_view(lgr_prep* p, int32_t i, lgr_csc* view);   /* m x n_i, f32 values, device pointers */
LGR_API int lgr_prep_scaled_download(lgr_prep* p, int32_t i, int32_t* colptr, int32_t* rowidx, double* values);
LGR_API void lgr_prep_destroy(lgr_prep* p);

/* ---- online iNMF (front end R/integration.R:722-933; the arithmetic is RcppPlanc::onlineINMF, R/integration.R:902-910,
   which is not part of the reference tree: restated from the published algorithm -- Gao et al. 2021, Algorithm 1 -- with
   the bookkeeping documented in DESIGN.md section 8; PARITY UNPINNED).  k <= 32, d <= 16.
   Scenario 1 (n_prev = 0, flags = 0): all d datasets from scratch.
   Scenario 2 (newDatasets, R/integration.R:755-812; n_prev > 0): datasets [0, n_prev) are the already factorised ones --
   their matrices are not read; Winit, Vinit[i], Ainit[i], Binit[i] carry their state, which enters the W updates and is
   returned unchanged -- and the minibatches run over datasets [n_prev, d) only.
   Scenario 3 (projection = TRUE; LGR_ONLINE_PROJECT): no update at all; H_i of the d given datasets against Winit alone
   (V = 0).  Only out->H (and objErr = sum ||X_i - W H_i||^2) is written. ---- */
#define LGR_ONLINE_PROJECT 1u
typedef struct lgr_online_args {
    int32_t d;               /* datasets */
    int32_t k;
    int64_t m;               /* shared genes */
    const lgr_csc* E;        /* d host matrices, m x n_i (scaleData) */
    double lambda;
    int32_t max_epochs;      /* [5] */
    int32_t minibatch_size;  /* [5000] cells per minibatch over all datasets; dataset i contributes round(n_i / N * size) */
    int32_t hals_iter;       /* [1] HALS sweeps per minibatch */
    uint32_t seed;           /* [1] set.seed(): W draw, the cells V_i starts from, the per-epoch permutations */
    const double* Winit;     /* m x k or NULL */
    const double* const* Vinit; /* d x (m x k) or NULL (scenario 1: both or neither; scenario 2: entries [0, n_prev) required,
                                   the others may be NULL = drawn) */
    int32_t device;
    uint32_t flags;          /* 0 or LGR_ONLINE_PROJECT */
    int32_t n_prev;          /* scenario 2: leading datasets that are already factorised (their E entries are ignored) */
    int32_t pad1;
    const double* const* Ainit; /* d x (k x k); entries [0, n_prev) are read (scenario 2), else NULL */
    const double* const* Binit; /* d x (m x k); entries [0, n_prev) are read (scenario 2), else NULL */
} lgr_online_args;
typedef struct lgr_online_out {
    double* W;               /* m x k */
    double** V;              /* d x (m x k) */
    double** H;              /* d x (n_i x k), cell x factor */
    double** A;              /* d x (k x k) or NULL */
    double** B;              /* d x (m x k) or NULL */
    double objErr;
    int32_t iterations;      /* minibatch iterations that were run */
    int32_t pad0;
    double loop_ms;          /* device time of the minibatch loop + final H solve (CUDA events) */
} lgr_online_out;
LGR_API int lgr_online_inmf(const lgr_online_args* args, lgr_online_out* out);

/* ---- online iNMF over matrices that are not in host memory as a whole.  The reference streams them from disk: scaleData kept
   in an HDF5 group is wrapped as RcppPlanc's H5SpMat (R/integration.R:745-750, .H5GroupToH5SpMat R/h5Utility.R:356-363) and
   read chunk by chunk.  Here the caller supplies a reader of column chunks; the library pulls every dataset through it once
   (chunk_cols columns at a time) into its resident fp32 copy in HBM -- 180 GB hold ~2e10 non-zeros, so the device is the
   cache the on-disk path lacks -- and then runs exactly lgr_online_inmf's loop.  args->E[i] carries only nrow / ncol for the
   datasets that are read (pointers ignored).  The reader fills, for columns [col0, col0 + ncol) of `dataset`: colptr
   (ncol + 1 entries, starting at 0), rowidx and values (at most `capacity` entries) and returns the chunk's number of
   non-zeros; if the chunk does not fit it returns -(needed capacity) and is called again; any other negative value aborts
   (LGR_ERR_INVALID).  It is also asked for single columns (the k cells a new V_i starts from). ---- */
typedef int64_t (*lgr_read_cols_fn)(void* user, int32_t dataset, int64_t col0, int64_t ncol, int32_t* colptr, int32_t* rowidx,
                                    double* values, int64_t capacity);
typedef struct lgr_chunk_source {
    lgr_read_cols_fn read_cols;
    void* user;
    const int64_t* nnz;      /* d entries: non-zeros of every dataset (HDF5: the length of its data array) */
    int64_t chunk_cols;      /* [1000] columns per call */
    int64_t max_chunk_nnz;   /* first guess of the capacity (the reader can ask for more) */
} lgr_chunk_source;
LGR_API int lgr_online_inmf_chunked(const lgr_online_args* args, const lgr_chunk_source* src, lgr_online_out* out);

/* ---- factor alignment: quantileNorm (R/integration.R:1621-1711 .quantileNorm.HList; src/quantile_norm.cpp:8-67
   max_factor_rcpp / cluster_vote_rcpp; RANN::nn2 = ANN 1.1.2 kd-tree search; stats::quantile type 7; stats::approxfun
   rule = 2; base::sample) -- the consumer of the H_i, run where they live.  Defaults of the reference in brackets. ---- */
typedef struct lgr_qn_params {
    int32_t quantiles;      /* [50]  number of quantile intervals: knots at seq(0, 1, by = 1 / quantiles) */
    int32_t reference;      /* 0-based index of the reference dataset */
    int32_t min_cells;      /* [20]  clusters smaller than this on either side are left untouched */
    int32_t n_neighbors;    /* [20]  k of the kNN refinement */
    int32_t center;         /* [0]   centre the columns before the arg-max */
    int32_t max_sample;     /* [1000] cells sampled per cluster for the quantiles (2..4096) */
    int32_t refine_knn;     /* [1]   majority vote over the approximate kNN graph */
    uint32_t seed;          /* [1]   set.seed() of the sample() stream */
    double eps;             /* [0.9] error bound of RANN::nn2 */
    int32_t n_dims_use;     /* 0 = all factors */
    int32_t pad0;
    const int32_t* dims_use; /* 1-based factor ids (useDims), or NULL */
} lgr_qn_params;
typedef struct lgr_qn_info {
    double cluster_ms, knn_ms, warp_ms;   /* device + host time of the three stages (CUDA events) */
    double tree_ms, search_ms;            /* inside knn_ms: host kd-tree builds (wall clock), device searches (CUDA events) */
    int32_t vote_sweeps;                  /* sweeps the in-place vote needed to reach its fixed point (max over datasets) */
    int32_t warped_clusters;              /* (dataset, cluster) pairs that were warped */
    int32_t sampled;                      /* 1 if some cluster exceeded max_sample (R's sample() stream was replayed) */
    int32_t pad0;
} lgr_qn_info;
/* H[i]: n_i x k column-major (cell x factor, i.e. t(H) of the liger object); Hnorm[i]: same shape (rbind them for H.norm);
   clusters[i]: n_i labels = 1-based factor id, -1 where no factor is positive.  info may be NULL. */
LGR_API int lgr_quantile_norm(int32_t d, int32_t k, const int64_t* n, const double* const* H, const lgr_qn_params* params,
                              double* const* Hnorm, int32_t* const* clusters, lgr_qn_info* info, int32_t device);

/* ---- factor alignment: centroidAlign (R/integration.R:2015-2084 .centroidAlign.Hraw; src/data_processing.cpp:10-20,61-85
   normalize_byCol_dense_rcpp / safe_scale; src/centoidAlign.cpp:16-85 moe_correct_ridge_cpp).  Defaults in brackets. ---- */
typedef struct lgr_ca_params {
    const double* lambda;    /* [1 each] ridge penalty per dataset (d values) */
    int32_t scale_emb;       /* [1] */
    int32_t center_emb;      /* [1] */
    int32_t scale_cluster;   /* [0] */
    int32_t center_cluster;  /* [0] */
    int32_t shift;           /* [0] subtract the global minimum of the memberships before normalising them */
    int32_t n_dims_use;      /* 0 = all factors */
    const int32_t* dims_use; /* 1-based factor ids (useDims), or NULL */
} lgr_ca_params;
/* H[i]: n_i x k column-major; aligned[i]: n_i x kd (kd = number of used dims; rbind them for H.norm); the three diagnostics
   (diagnosis = TRUE: which.max of the raw row, of the aligned row, of the membership column; 1-based) may be NULL. */
LGR_API int lgr_centroid_align(int32_t d, int32_t k, const int64_t* n, const double* const* H, const lgr_ca_params* params,
                               double* const* aligned, int32_t* const* raw_which_max, int32_t* const* z_which_max,
                               int32_t* const* r_which_max, int32_t device);

/* ---- resident-handle API (data stays in HBM between calls; used by bench + multi-GPU hosts) ---- */
typedef struct lgr_ctx lgr_ctx;
LGR_API int lgr_ctx_create(const lgr_inmf_args* args, lgr_ctx** ctx); /* uploads E/P, builds device layouts, sets inits */
LGR_API int lgr_ctx_reset(lgr_ctx* ctx, const double* Winit, const double* const* Vinit, const double* const* Uinit);
LGR_API int lgr_ctx_iterate(lgr_ctx* ctx, int32_t niter, double* obj_traj /* NULL or niter */);
/* niter iterations followed by the final objective -- what lgr_inmf runs between its upload and its download -- inside
   ONE CUDA-event bracket (lgr_timings.loop_ms then covers the loop and the objective pass). */
LGR_API int lgr_ctx_run(lgr_ctx* ctx, int32_t niter, double* objErr);
LGR_API int lgr_ctx_objective(lgr_ctx* ctx, double* obj);
/* The three block solves of the R-level ANLS (inmf_solveH / inmf_solveV / inmf_solveW, R/cINMF.R:477-503; their callers
   R/cINMF.R:392-407, R/optimizeNewParam.R:97-134,288-291,327-329) on the RESIDENT data: X is uploaded once by
   lgr_ctx_create, the right-hand sides are formed by the same sweeps as in lgr_inmf, nothing but the factors a caller asks
   for crosses PCIe.  `blocks` is any OR of LGR_SOLVE_*; they run in the order H -> V -> W, each from the factors that are on
   the device when it starts (H: CtC = (W+V)^T(W+V) + lambda V^T V, CtB = (W+V)^T E; V: CtC = (1+lambda) H^T H,
   CtB = H^T E^T - H^T H W^T; W: CtC = sum H^T H, CtB = sum (H^T E^T - H^T H V^T)). */
#define LGR_SOLVE_H 1u
#define LGR_SOLVE_V 2u
#define LGR_SOLVE_W 4u
LGR_API int lgr_ctx_solve(lgr_ctx* ctx, uint32_t blocks);
/* replace resident factors: W m x k, V[i] m x k, U[i] u_i x k, H[i] n_i x k, column-major; NULL (array or entry) = keep */
LGR_API int lgr_ctx_set_factors(lgr_ctx* ctx, const double* W, const double* const* V, const double* const* U,
                                const double* const* H);
/* the two products over X for callers that assemble their own normal equations (residual right-hand sides of
   optimizeNewK / optimizeNewData): which = 0: out[i] (n_i x k) = [E_i; P_i]^T ([W; 0] + [V_i; U_i]);
   which = 1: out[i] ((m + u_i) x k) = [E_i; P_i] H_i.  Column-major; NULL entries are skipped. */
LGR_API int lgr_ctx_products(lgr_ctx* ctx, int32_t which, double* const* out);
/* objErr_i (src/objErr.cpp:13-31) of dataset i from the resident factors -- no upload, no layout build */
LGR_API int lgr_ctx_objerr_i(lgr_ctx* ctx, int32_t i, double* obj);
LGR_API int lgr_ctx_download(lgr_ctx* ctx, lgr_inmf_out* out);
/* quantileNorm of the RESIDENT H_i (no upload of H; single-rank contexts only) */
/* centroidAlign of the RESIDENT H_i (single-rank contexts only) */
LGR_API int lgr_ctx_centroid_align(lgr_ctx* ctx, const lgr_ca_params* params, double* const* aligned);
LGR_API int lgr_ctx_quantile_norm(lgr_ctx* ctx, const lgr_qn_params* params, double* const* Hnorm, int32_t* const* clusters,
                                  lgr_qn_info* info);
LGR_API int lgr_ctx_timings(lgr_ctx* ctx, lgr_timings* t);
LGR_API void* lgr_ctx_stream(lgr_ctx* ctx); /* the cudaStream_t all kernels are launched on */
/* per-kernel device time of the last lgr_ctx_iterate (CUDA events on the launch stream):
   names[i] / ms[i] / launches[i] for i < *n (n in: capacity, out: count). */
LGR_API int lgr_ctx_kernel_times(lgr_ctx* ctx, int32_t* n, const char** names, double* ms, int64_t* launches);
LGR_API int lgr_ctx_set_kernel_timing(lgr_ctx* ctx, int32_t on); /* record CUDA events around every kernel of the loop */
LGR_API void lgr_ctx_destroy(lgr_ctx* ctx);

/* ---- seeded initialisation: R's Mersenne-Twister stream (R/integration.R:412-437) ---------- */
typedef struct lgr_rng lgr_rng;
LGR_API lgr_rng* lgr_r_set_seed(uint32_t seed);
LGR_API void lgr_r_runif(lgr_rng* rng, int64_t n, double lo, double hi, double* out); /* abs() applied by caller */
LGR_API void lgr_r_free(lgr_rng* rng);

/* ---- misc ---------------------------------------------------------------------------------- */
LGR_API const char* lgr_last_error(void);
LGR_API int lgr_device_count(void);
LGR_API int lgr_version(void);                      /* major * 1000 + minor */
LGR_API int lgr_nccl_unique_id(void* out128);       /* rank 0 calls this and ships the 128 bytes to the others */
/* page-locked host memory for callers that want full-speed PCIe copies of their inputs / outputs (optional) */
LGR_API void* lgr_host_alloc(size_t bytes);
LGR_API void lgr_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* LIGER_B200_H */
