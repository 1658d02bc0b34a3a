"""ctypes binding of include/liger_b200.h (the C ABI of the CUDA library).

There is no Python or CPU implementation behind these calls: if the shared library is
missing, or no CUDA device is present, the calls fail loudly."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libliger_b200.so")

LGR_F64, LGR_F32 = 0, 1
FLAG_TRAJECTORY, FLAG_COLD_START, FLAG_PINNED_IO, FLAG_LOCAL_SHARD, FLAG_TENSOR_U, FLAG_LEGACY_NNLS = 0x1, 0x2, 0x4, 0x8, 0x10, 0x20
FLAG_NO_DENSE = 0x40
FLAG_DATASET_SHARD = 0x80
SOLVE_H, SOLVE_V, SOLVE_W = 1, 2, 4
STATUS = {0: "OK", 1: "INVALID", 2: "NO_DEVICE", 3: "CUDA", 4: "NCCL", 5: "ALLOC", 6: "INTERRUPT"}

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_uint64_p = C.POINTER(C.c_uint64)


class lgr_csc(C.Structure):
    _fields_ = [("nrow", C.c_int64), ("ncol", C.c_int64), ("colptr", c_int32_p), ("rowidx", c_int32_p),
                ("values", C.c_void_p), ("value_dtype", C.c_int32), ("location", C.c_int32),
                ("row_scale", c_double_p), ("col_scale", c_double_p)]


INTERRUPT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p)


class lgr_inmf_args(C.Structure):
    _fields_ = [("d", C.c_int32), ("k", C.c_int32), ("niter", C.c_int32), ("flags", C.c_uint32),
                ("m", C.c_int64), ("E", C.POINTER(lgr_csc)), ("P", C.POINTER(lgr_csc)),
                ("lambda_", c_double_p), ("Winit", c_double_p), ("Vinit", C.POINTER(c_double_p)),
                ("Uinit", C.POINTER(c_double_p)), ("Hinit", C.POINTER(c_double_p)),
                ("device", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32), ("init_seed", C.c_int32),
                ("nccl_unique_id", C.c_void_p), ("interrupt", INTERRUPT_FN), ("interrupt_user", C.c_void_p)]


class lgr_timings(C.Structure):
    _fields_ = [("upload_ms", C.c_double), ("loop_ms", C.c_double), ("download_ms", C.c_double),
                ("h_phase_ms", C.c_double), ("vw_phase_ms", C.c_double), ("comm_ms", C.c_double),
                ("kernel_launches", C.c_int64), ("nnls_unconverged", C.c_int64), ("nnls_pivots", C.c_int64)]


class lgr_inmf_out(C.Structure):
    _fields_ = [("W", c_double_p), ("V", C.POINTER(c_double_p)), ("U", C.POINTER(c_double_p)),
                ("H", C.POINTER(c_double_p)), ("H_passive", C.POINTER(c_uint64_p)), ("obj_traj", c_double_p),
                ("objErr", C.c_double), ("timings", lgr_timings)]


class lgr_qn_params(C.Structure):
    _fields_ = [("quantiles", C.c_int32), ("reference", C.c_int32), ("min_cells", C.c_int32), ("n_neighbors", C.c_int32),
                ("center", C.c_int32), ("max_sample", C.c_int32), ("refine_knn", C.c_int32), ("seed", C.c_uint32),
                ("eps", C.c_double), ("n_dims_use", C.c_int32), ("pad0", C.c_int32), ("dims_use", C.POINTER(C.c_int32))]


class lgr_online_args(C.Structure):
    _fields_ = [("d", C.c_int32), ("k", C.c_int32), ("m", C.c_int64), ("E", C.c_void_p), ("lambda_", C.c_double),
                ("max_epochs", C.c_int32), ("minibatch_size", C.c_int32), ("hals_iter", C.c_int32), ("seed", C.c_uint32),
                ("Winit", c_double_p), ("Vinit", C.POINTER(c_double_p)), ("device", C.c_int32), ("flags", C.c_uint32),
                ("n_prev", C.c_int32), ("pad1", C.c_int32), ("Ainit", C.POINTER(c_double_p)), ("Binit", C.POINTER(c_double_p))]


READ_COLS_FN = C.CFUNCTYPE(C.c_int64, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                           C.POINTER(C.c_double), C.c_int64)


class lgr_chunk_source(C.Structure):
    _fields_ = [("read_cols", READ_COLS_FN), ("user", C.c_void_p), ("nnz", C.POINTER(C.c_int64)), ("chunk_cols", C.c_int64),
                ("max_chunk_nnz", C.c_int64)]


class lgr_online_out(C.Structure):
    _fields_ = [("W", c_double_p), ("V", C.POINTER(c_double_p)), ("H", C.POINTER(c_double_p)), ("A", C.POINTER(c_double_p)),
                ("B", C.POINTER(c_double_p)), ("objErr", C.c_double), ("iterations", C.c_int32), ("pad0", C.c_int32),
                ("loop_ms", C.c_double)]


class lgr_ca_params(C.Structure):
    _fields_ = [("lambda_", c_double_p), ("scale_emb", C.c_int32), ("center_emb", C.c_int32), ("scale_cluster", C.c_int32),
                ("center_cluster", C.c_int32), ("shift", C.c_int32), ("n_dims_use", C.c_int32),
                ("dims_use", C.POINTER(C.c_int32))]


class lgr_qn_info(C.Structure):
    _fields_ = [("cluster_ms", C.c_double), ("knn_ms", C.c_double), ("warp_ms", C.c_double), ("tree_ms", C.c_double),
                ("search_ms", C.c_double), ("vote_sweeps", C.c_int32),
                ("warped_clusters", C.c_int32), ("sampled", C.c_int32), ("pad0", C.c_int32)]


class LigerB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"liger_b200 [{STATUS.get(code, code)}]: {msg}")
        self.code = code


_lib = None

# every symbol include/liger_b200.h declares (tests check that the .so exports all of them)
EXPORTS = [
    "lgr_inmf", "lgr_uinmf", "lgr_bppnnls_prod", "lgr_objerr_i", "lgr_ctx_create", "lgr_ctx_reset",
    "lgr_ctx_iterate", "lgr_ctx_run", "lgr_ctx_objective", "lgr_ctx_solve", "lgr_ctx_set_factors", "lgr_ctx_products",
    "lgr_ctx_objerr_i", "lgr_quantile_norm", "lgr_ctx_quantile_norm", "lgr_centroid_align", "lgr_ctx_centroid_align", "lgr_online_inmf", "lgr_online_inmf_chunked", "lgr_ctx_download", "lgr_ctx_timings", "lgr_ctx_stream",
    "lgr_ctx_kernel_times", "lgr_ctx_set_kernel_timing", "lgr_ctx_destroy", "lgr_r_set_seed", "lgr_r_runif",
    "lgr_r_free", "lgr_last_error", "lgr_device_count", "lgr_version", "lgr_nccl_unique_id", "lgr_host_alloc",
    "lgr_host_free", "lgr_normalize_csc", "lgr_row_vars_sparse", "lgr_scale_not_center_by_row",
    "lgr_prep_create", "lgr_prep_normalize", "lgr_prep_gene_stats", "lgr_prep_scale", "lgr_prep_scaled_view",
    "lgr_prep_scaled_download", "lgr_prep_destroy",
]


def lib():
    """Load libliger_b200.so (built in-tree by `python -m liger_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LigerB200Error(2, f"{LIB_PATH} is missing: build it with `python -m liger_b200.build` "
                                "(there is no Python/CPU fallback for the iNMF hot path)")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    L.lgr_last_error.restype = C.c_char_p
    L.lgr_inmf.argtypes = [C.POINTER(lgr_inmf_args), C.POINTER(lgr_inmf_out)]
    L.lgr_uinmf.argtypes = [C.POINTER(lgr_inmf_args), C.POINTER(lgr_inmf_out)]
    L.lgr_bppnnls_prod.argtypes = [C.c_int32, C.c_int64, c_double_p, c_double_p, c_double_p, C.c_int32]
    L.lgr_objerr_i.argtypes = [C.c_int32, c_double_p, c_double_p, c_double_p, C.POINTER(lgr_csc), C.c_double,
                               C.c_int32, c_double_p]
    L.lgr_ctx_create.argtypes = [C.POINTER(lgr_inmf_args), C.POINTER(C.c_void_p)]
    L.lgr_ctx_reset.argtypes = [C.c_void_p, c_double_p, C.POINTER(c_double_p), C.POINTER(c_double_p)]
    L.lgr_ctx_iterate.argtypes = [C.c_void_p, C.c_int32, c_double_p]
    L.lgr_ctx_run.argtypes = [C.c_void_p, C.c_int32, c_double_p]
    L.lgr_ctx_objective.argtypes = [C.c_void_p, c_double_p]
    L.lgr_ctx_solve.argtypes = [C.c_void_p, C.c_uint32]
    L.lgr_ctx_set_factors.argtypes = [C.c_void_p, c_double_p, C.POINTER(c_double_p), C.POINTER(c_double_p), C.POINTER(c_double_p)]
    L.lgr_ctx_products.argtypes = [C.c_void_p, C.c_int32, C.POINTER(c_double_p)]
    L.lgr_ctx_objerr_i.argtypes = [C.c_void_p, C.c_int32, c_double_p]
    L.lgr_ctx_download.argtypes = [C.c_void_p, C.POINTER(lgr_inmf_out)]
    L.lgr_quantile_norm.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(c_double_p), C.POINTER(lgr_qn_params),
                                    C.POINTER(c_double_p), C.POINTER(C.POINTER(C.c_int32)), C.POINTER(lgr_qn_info), C.c_int32]
    L.lgr_centroid_align.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(c_double_p), C.POINTER(lgr_ca_params),
                                     C.POINTER(c_double_p), C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.POINTER(C.c_int32)),
                                     C.POINTER(C.POINTER(C.c_int32)), C.c_int32]
    L.lgr_online_inmf.argtypes = [C.POINTER(lgr_online_args), C.POINTER(lgr_online_out)]
    L.lgr_online_inmf_chunked.argtypes = [C.POINTER(lgr_online_args), C.POINTER(lgr_chunk_source), C.POINTER(lgr_online_out)]
    L.lgr_ctx_centroid_align.argtypes = [C.c_void_p, C.POINTER(lgr_ca_params), C.POINTER(c_double_p)]
    L.lgr_ctx_quantile_norm.argtypes = [C.c_void_p, C.POINTER(lgr_qn_params), C.POINTER(c_double_p),
                                        C.POINTER(C.POINTER(C.c_int32)), C.POINTER(lgr_qn_info)]
    L.lgr_ctx_timings.argtypes = [C.c_void_p, C.POINTER(lgr_timings)]
    L.lgr_ctx_stream.argtypes = [C.c_void_p]
    L.lgr_ctx_stream.restype = C.c_void_p
    L.lgr_ctx_kernel_times.argtypes = [C.c_void_p, c_int32_p, C.POINTER(C.c_char_p), c_double_p,
                                       C.POINTER(C.c_int64)]
    L.lgr_ctx_set_kernel_timing.argtypes = [C.c_void_p, C.c_int32]
    L.lgr_ctx_destroy.argtypes = [C.c_void_p]
    L.lgr_ctx_destroy.restype = None
    L.lgr_r_set_seed.argtypes = [C.c_uint32]
    L.lgr_r_set_seed.restype = C.c_void_p
    L.lgr_r_runif.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_double, c_double_p]
    L.lgr_r_runif.restype = None
    L.lgr_r_free.argtypes = [C.c_void_p]
    L.lgr_r_free.restype = None
    L.lgr_normalize_csc.argtypes = [C.c_int64, c_int32_p, c_double_p, c_double_p, c_double_p, C.c_int32]
    L.lgr_row_vars_sparse.argtypes = [C.c_int64, C.c_int64, c_int32_p, c_int32_p, c_double_p, c_double_p, C.c_double,
                                      c_double_p, c_double_p, C.c_int32]
    L.lgr_scale_not_center_by_row.argtypes = [C.c_int64, C.c_int64, c_int32_p, c_int32_p, c_double_p, c_double_p,
                                              c_double_p, C.c_int32]
    L.lgr_prep_create.argtypes = [C.c_int32, C.POINTER(lgr_csc), C.c_int32, C.POINTER(C.c_void_p)]
    L.lgr_prep_normalize.argtypes = [C.c_void_p]
    L.lgr_prep_gene_stats.argtypes = [C.c_void_p, C.c_int32, c_double_p, c_double_p]
    L.lgr_prep_scale.argtypes = [C.c_void_p, C.c_int32, C.c_int64, c_int32_p]
    L.lgr_prep_scaled_view.argtypes = [C.c_void_p, C.c_int32, C.POINTER(lgr_csc)]
    L.lgr_prep_scaled_download.argtypes = [C.c_void_p, C.c_int32, c_int32_p, c_int32_p, c_double_p]
    L.lgr_prep_destroy.argtypes = [C.c_void_p]
    L.lgr_prep_destroy.restype = None
    L.lgr_nccl_unique_id.argtypes = [C.c_void_p]
    L.lgr_host_alloc.argtypes = [C.c_size_t]
    L.lgr_host_alloc.restype = C.c_void_p
    L.lgr_host_free.argtypes = [C.c_void_p]
    L.lgr_host_free.restype = None
    _lib = L
    return L


def check(code):
    if code != 0:
        raise LigerB200Error(code, lib().lgr_last_error().decode("utf-8", "replace"))
