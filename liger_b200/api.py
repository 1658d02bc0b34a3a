"""Host-side mirror of the native calls rliger makes on the iNMF / UINMF path.

Function names, argument names/meaning and return shapes follow the reference's call sites
(welch-lab/liger):

    inmf(objectList, k, lambda_, niter, Hinit, Vinit, Winit, nCores, verbose)
        <- RcppPlanc::inmf            R/integration.R:438-441
    uinmf(objectList, unsharedList, k, lambda_, niter, nCores, verbose)
        <- RcppPlanc::uinmf           R/integration.R:1285-1287
    bppnnls_prod(CtC, CtB, nCores)    <- RcppPlanc::bppnnls_prod   R/cINMF.R:481,489,501
    objErr_i(H, W, V, E, lambda_)     <- objErr_i                  src/objErr.cpp:13-31

Everything is computed by the CUDA library behind include/liger_b200.h; this module only
marshals numpy / scipy buffers (the ctypes analogue of the Rcpp shim in INTEGRATION.md).
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import _ffi
from ._ffi import lgr_csc, lgr_inmf_args, lgr_inmf_out, c_double_p, c_int32_p, c_uint64_p


def _f64(a, order="F"):
    return np.require(np.asarray(a, dtype=np.float64), dtype=np.float64, requirements=[order.upper() + "_CONTIGUOUS", "ALIGNED"])


def _dp(a):
    return a.ctypes.data_as(c_double_p)


class _Pinned:
    """Owner of one page-locked host allocation (lgr_host_alloc); numpy arrays are views of it.  Freed blocks go
    back to a small cache (cudaHostAlloc costs ~0.4 ms per MB), mirroring the library's device memory pool."""
    _cache = {}
    _cached_bytes = 0
    _CACHE_LIMIT = 4 << 30

    def __init__(self, nbytes):
        self.nbytes = max(int(nbytes), 1)
        free = _Pinned._cache.get(self.nbytes)
        if free:
            self.ptr = free.pop()
            _Pinned._cached_bytes -= self.nbytes
        else:
            self.ptr = _ffi.lib().lgr_host_alloc(self.nbytes)
            if not self.ptr:
                raise MemoryError("lgr_host_alloc failed")

    def __del__(self):
        try:
            if _Pinned._cached_bytes + self.nbytes <= _Pinned._CACHE_LIMIT:
                _Pinned._cache.setdefault(self.nbytes, []).append(self.ptr)
                _Pinned._cached_bytes += self.nbytes
            else:
                _ffi.lib().lgr_host_free(self.ptr)
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float64):
    """Fortran-ordered array in page-locked host memory (full-speed D2H for the outputs); NOT zero-initialised."""
    n = int(np.prod(shape))
    own = _Pinned(n * np.dtype(dtype).itemsize)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(own.ptr)
    buf._lgr_owner = own        # the ctypes buffer (base of the array) keeps the allocation alive
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape, order="F")


class DeviceCsc:
    """A genes x cells CSC matrix that lives in HBM (a view handed out by lgr_prep_scaled_view).  It can be passed to
    inmf / uinmf / Context in place of a scipy matrix; `owner` keeps the preprocessing handle alive."""

    def __init__(self, struct, owner):
        self.struct = struct
        self.owner = owner
        self.shape = (int(struct.nrow), int(struct.ncol))


class CountScaled:
    """scaleData together with its count structure  X[g, c] = count[g, c] * row_scale[g] * col_scale[c]  (integer counts
    in 1..256; row_scale = 1 / row rms of scaleNotCenter, col_scale = 1 / library size of normalize).  The library verifies
    the structure on upload and then runs the two sweeps over X on the tensor cores (csrc/lgr_dense.cu)."""

    def __init__(self, mat, row_scale, col_scale):
        self.mat = mat
        self.row_scale = np.ascontiguousarray(row_scale, dtype=np.float64)
        self.col_scale = np.ascontiguousarray(col_scale, dtype=np.float64)
        self.shape = mat.shape if hasattr(mat, "shape") else tuple(mat[3])
        if self.row_scale.shape != (self.shape[0],) or self.col_scale.shape != (self.shape[1],):
            raise ValueError("row_scale / col_scale must have one entry per row / column")


class _Csc:
    """Keeps the numpy buffers of one dgCMatrix-like input alive and exposes an lgr_csc view."""

    def __init__(self, mat, keep_f32=False):
        if isinstance(mat, DeviceCsc):   # resident scaleData (preprocess.Pipeline): device pointers, nothing to copy
            self.owner = mat
            self.shape = mat.shape
            self.struct = mat.struct
            return
        scales = None
        if isinstance(mat, CountScaled):
            scales = mat
            mat = mat.mat
        if sp.issparse(mat):
            m = mat.tocsc()
            indptr, indices, data, shape = m.indptr, m.indices, m.data, m.shape
        else:  # (indptr, indices, data, shape) tuple, e.g. pinned buffers prepared by bench.py
            indptr, indices, data, shape = mat
        if int(indptr[-1]) >= 2 ** 31:
            raise ValueError("more than 2^31-1 non-zeros in one matrix (dgCMatrix limit)")
        self.indptr = np.ascontiguousarray(indptr, dtype=np.int32)
        self.indices = np.ascontiguousarray(indices, dtype=np.int32)
        if data.dtype == np.float32 and keep_f32:
            self.data = np.ascontiguousarray(data)
            dt = _ffi.LGR_F32
        else:
            self.data = np.ascontiguousarray(data, dtype=np.float64)
            dt = _ffi.LGR_F64
        self.shape = (int(shape[0]), int(shape[1]))
        self.struct = lgr_csc(self.shape[0], self.shape[1], self.indptr.ctypes.data_as(c_int32_p),
                              self.indices.ctypes.data_as(c_int32_p), self.data.ctypes.data_as(C.c_void_p), dt, 0)
        if scales is not None:
            self.scales = scales
            self.struct.row_scale = _dp(scales.row_scale)
            self.struct.col_scale = _dp(scales.col_scale)


class _Args:
    """Builds lgr_inmf_args and owns every buffer it points to."""

    def __init__(self, objectList, k, lambda_, niter, Winit, Vinit, Hinit=None, unsharedList=None, Uinit=None,
                 flags=0, device=-1, rank=0, world=1, nccl_id=None, keep_f32=False, init_seed=1):
        d = len(objectList)
        if d < 1:
            raise ValueError("objectList is empty")
        self.E = [_Csc(x, keep_f32) for x in objectList]
        self.m = self.E[0].shape[0]
        self.k = int(k)
        self.d = d
        self.n = [e.shape[1] for e in self.E]
        self.u = [0] * d
        self.Earr = (lgr_csc * d)(*[e.struct for e in self.E])
        self.Parr = None
        self.P = [None] * d
        if unsharedList is not None:
            ps = []
            for i, p in enumerate(unsharedList):
                if p is None or (p.shape[0] if hasattr(p, "shape") else p[3][0]) == 0:
                    ps.append(lgr_csc(0, self.n[i], None, None, None, 0, 0))
                else:
                    self.P[i] = _Csc(p, keep_f32)
                    self.u[i] = self.P[i].shape[0]
                    ps.append(self.P[i].struct)
            self.Parr = (lgr_csc * d)(*ps)
        lam = np.broadcast_to(np.asarray(lambda_, dtype=np.float64), (d,)).copy()
        self.lam = lam
        if (Winit is None) != (Vinit is None):
            raise ValueError("Winit and Vinit must both be given, or both be None (in-library seeded initialisation)")
        self.W0 = self.V0 = self.Varr = None
        if Winit is not None:
            self.W0 = _f64(Winit)
            if self.W0.shape != (self.m, self.k):
                raise ValueError(f"Winit must be {self.m} x {self.k}")
            self.V0 = [_f64(v) for v in Vinit]
            for v in self.V0:
                if v.shape != (self.m, self.k):
                    raise ValueError(f"every Vinit must be {self.m} x {self.k}")
            self.Varr = (c_double_p * d)(*[_dp(v) for v in self.V0])
        self.Uarr = None
        if unsharedList is not None and Winit is not None:
            self.U0 = [None if self.u[i] == 0 else _f64(Uinit[i]) for i in range(d)]
            self.Uarr = (c_double_p * d)(*[(_dp(x) if x is not None else None) for x in self.U0])
        self.Harr = None
        if Hinit is not None:
            self.H0 = [_f64(h) for h in Hinit]
            for h, n in zip(self.H0, self.n):
                if h.shape != (n, self.k):
                    raise ValueError("every Hinit must be n_i x k")
            self.Harr = (c_double_p * d)(*[_dp(h) for h in self.H0])
        self.nccl_id = nccl_id
        a = lgr_inmf_args()
        a.d, a.k, a.niter, a.flags, a.m = d, self.k, int(niter), int(flags), self.m
        a.E = self.Earr
        a.P = self.Parr
        a.lambda_ = _dp(lam)
        a.Winit = _dp(self.W0) if self.W0 is not None else None
        a.Vinit = self.Varr
        a.init_seed = int(init_seed)
        a.Uinit = self.Uarr
        a.Hinit = self.Harr
        a.device, a.rank, a.world = int(device), int(rank), int(world)
        if nccl_id is not None:
            self._idbuf = C.create_string_buffer(bytes(nccl_id), 128)
            a.nccl_unique_id = C.cast(self._idbuf, C.c_void_p)
        self.struct = a


class _Out:
    def __init__(self, args, want_passive=False, traj=False, pinned=False, reuse=None):
        d, k, m = args.d, args.k, args.m
        if reuse is not None:      # caller-owned result buffers (a host that keeps page-locked buffers between calls)
            self.W, self.V, self.H = reuse["W"], list(reuse["V"]), list(reuse["H"])
            ok = (self.W.shape == (m, k) and len(self.V) == d and len(self.H) == d and
                  all(v.shape == (m, k) for v in self.V) and all(h.shape == (n, k) for h, n in zip(self.H, args.n)) and
                  all(x.dtype == np.float64 and x.flags.f_contiguous for x in [self.W] + self.V + self.H))
            if not ok:
                raise ValueError("outputs: W (m x k), V (d x m x k), H (d x n_i x k) as Fortran-ordered float64 arrays")
        else:
            self.W = np.zeros((m, k), order="F")
            self.V = [np.zeros((m, k), order="F") for _ in range(d)]
        if reuse is not None:
            pass
        elif pinned:
            self.H = [pinned_empty((n, k)) for n in args.n]
            if args.struct.world > 1 and not (args.struct.flags & _ffi.FLAG_LOCAL_SHARD):
                for h in self.H:      # slice mode: the library writes only this rank's rows
                    h[...] = 0.0
        else:
            self.H = [np.zeros((n, k), order="F") for n in args.n]
        self.U = [np.zeros((u, k), order="F") if u > 0 else None for u in args.u]
        self.Varr = (c_double_p * d)(*[_dp(v) for v in self.V])
        self.Harr = (c_double_p * d)(*[_dp(h) for h in self.H])
        self.Uarr = (c_double_p * d)(*[(_dp(u) if u is not None else None) for u in self.U])
        o = lgr_inmf_out()
        o.W = _dp(self.W)
        o.V = self.Varr
        o.H = self.Harr
        o.U = self.Uarr
        self.passive = None
        if want_passive:
            self.passive = [np.zeros(n, dtype=np.uint64) for n in args.n]
            self.Parr = (c_uint64_p * d)(*[p.ctypes.data_as(c_uint64_p) for p in self.passive])
            o.H_passive = self.Parr
        self.traj = None
        if traj:
            self.traj = np.zeros(max(int(args.struct.niter), 1))
            o.obj_traj = _dp(self.traj)
        self.struct = o

    def timings(self):
        t = self.struct.timings
        return {f: getattr(t, f) for f, _ in _ffi.lgr_timings._fields_}


def inmf(objectList, k=20, lambda_=5.0, niter=30, Hinit=None, Vinit=None, Winit=None, nCores=2, verbose=False,
         trajectory=False, device=-1, return_passive=False, cold_start=False, pinned_outputs=False, tensor_u=False,
         rank=0, world=1, nccl_id=None, local_shard=False, legacy_nnls=False, seed=1, pinned_io=False, no_dense=False,
         outputs=None, dataset_shard=False):
    """Drop-in for RcppPlanc::inmf (reference call site R/integration.R:438-441).

    outputs: optional dict(W, V, H) of caller-owned Fortran-ordered float64 result buffers (e.g. page-locked, kept between
    calls); the returned dict then holds those same arrays.

    Multi-GPU (one process per GPU): pass `rank`, `world`, the `nccl_id` created by rank 0 (nccl_unique_id()) and
    `local_shard=True` when objectList already holds only this rank's cells; H is then this rank's rows.  `dataset_shard=True`:
    whole-dataset placement -- objectList / Vinit hold only this rank's datasets (LGR_FLAG_DATASET_SHARD).

    objectList: list of scipy CSC matrices (genes x cells; the `scaleData` of each dataset).
    Hinit: list of n_i x k, Vinit: list of m x k, Winit: m x k.  `nCores`/`verbose` are accepted for
    signature parity and ignored (the work runs on the GPU).
    Returns dict(H=[n_i x k], V=[m x k], W=m x k, objErr=float) exactly like RcppPlanc
    (H is cell x factor; the caller transposes, R/integration.R:453).
    """
    if (Winit is None) != (Vinit is None):
        raise ValueError("Winit and Vinit must both be given, or both be None")
    # Winit = Vinit = None (R/suggestK.R:78-88 passes NULL for all three): the library draws W, V_i from R's
    # Mersenne-Twister stream seeded with `seed`, in .runINMF.list's order
    flags = ((_ffi.FLAG_TRAJECTORY if trajectory else 0) | (_ffi.FLAG_COLD_START if cold_start else 0) |
             (_ffi.FLAG_TENSOR_U if tensor_u else 0) | (0x8 if (local_shard and world > 1) else 0) |
             (_ffi.FLAG_LEGACY_NNLS if legacy_nnls else 0) | (_ffi.FLAG_PINNED_IO if pinned_io else 0) |
             (_ffi.FLAG_NO_DENSE if no_dense else 0) | (_ffi.FLAG_DATASET_SHARD if (dataset_shard and world > 1) else 0))
    a = _Args(objectList, k, lambda_, niter, Winit, Vinit, Hinit, flags=flags, device=device, rank=rank, world=world,
              nccl_id=nccl_id, init_seed=seed)
    o = _Out(a, want_passive=return_passive, traj=trajectory, pinned=pinned_outputs, reuse=outputs)
    _ffi.check(_ffi.lib().lgr_inmf(C.byref(a.struct), C.byref(o.struct)))
    res = {"H": o.H, "V": o.V, "W": o.W, "objErr": float(o.struct.objErr), "timings": o.timings()}
    if trajectory:
        res["traj"] = o.traj[:niter].copy()
    if return_passive:
        res["H_passive"] = o.passive
    return res


def uinmf(objectList, unsharedList, k=20, lambda_=5.0, niter=30, nCores=2, verbose=False, Winit=None, Vinit=None,
          Uinit=None, trajectory=False, device=-1, seed=1):
    """Drop-in for RcppPlanc::uinmf (reference call site R/integration.R:1285-1287).

    unsharedList[i]: u_i x n_i CSC or None.  lambda_: scalar or per-dataset vector
    (vignettes/articles/STARmap_dropviz_vig.Rmd:72).  The reference passes no inits (RcppPlanc draws
    them internally -- parity unpinned); here explicit W/V/U inits are required at this level and
    liger_b200.integration.run_uinmf_list draws them from R's RNG stream.
    Returns dict(H, V, W, U, objErr)."""
    given = [x is not None for x in (Winit, Vinit, Uinit)]
    if any(given) and not all(given):
        raise ValueError("liger_b200.uinmf: pass all of Winit/Vinit/Uinit, or none (the library then draws them from R's "
                         "RNG stream seeded with `seed`, as RcppPlanc::uinmf draws its own)")
    flags = _ffi.FLAG_TRAJECTORY if trajectory else 0
    a = _Args(objectList, k, lambda_, niter, Winit, Vinit, None, unsharedList=unsharedList, Uinit=Uinit, flags=flags,
              device=device, init_seed=seed)
    o = _Out(a, traj=trajectory)
    _ffi.check(_ffi.lib().lgr_uinmf(C.byref(a.struct), C.byref(o.struct)))
    res = {"H": o.H, "V": o.V, "W": o.W, "U": o.U, "objErr": float(o.struct.objErr), "timings": o.timings()}
    if trajectory:
        res["traj"] = o.traj[:niter].copy()
    return res


def online_inmf(objectList, k=20, lambda_=5.0, maxEpochs=5, minibatchSize=5000, HALSiter=1, Winit=None, Vinit=None, seed=1,
                nCores=2, verbose=False, device=-1, newDatasets=None, projection=False, Hinit=None, Ainit=None, Binit=None,
                chunk_source=None):
    """RcppPlanc::onlineINMF (reference call site R/integration.R:902-910).  PARITY UNPINNED against RcppPlanc (its source is
    not in the reference tree): the algorithm is the published one, specified line by line in DESIGN.md section 8.
    Scenario 1 (newDatasets = None): online iNMF over `objectList` from scratch.
    Scenario 2 (newDatasets given): `objectList` are the already factorised datasets -- only their number matters, their
    state comes in as Winit / Vinit / Ainit / Binit (one entry per dataset of objectList) -- and the minibatches run over
    `newDatasets`; W and the new V_i, A_i, B_i, H_i are learnt, the old ones returned as given (H: Hinit, if passed).
    Scenario 3 (projection = True): H of `newDatasets` against Winit alone; returns only dict(H=[...]).
    `chunk_source` (the reference's on-disk path: scaleData as an H5SpMat read chunk by chunk, R/integration.R:745-750): a dict
    `{"read": f(dataset, col0, ncol) -> (colptr, rowidx, values), "shapes": [(m, n_i)], "nnz": [...], "chunk_cols": 1000}`
    describing the matrices that are learnt from (scenario 1: objectList, else newDatasets -- pass None entries there); the
    library pulls every dataset through `read` once into its HBM-resident copy (lgr_online_inmf_chunked).
    Returns dict(H=[n_i x k], V=[m x k], W, A=[k x k], B=[m x k], objErr) like the R function (H is cell x factor; the
    caller transposes, R/integration.R:924)."""
    if projection:
        if newDatasets is None or Winit is None:
            raise ValueError("projection needs newDatasets and Winit")
        prev, data = [], list(newDatasets)
    elif newDatasets is not None:
        prev, data = list(objectList), list(newDatasets)
        if Winit is None or any(x is None for x in (Vinit, Ainit, Binit)) or \
                any(len(x) < len(prev) or any(y is None for y in x[:len(prev)]) for x in (Vinit, Ainit, Binit)):
            raise ValueError("Cannot find complete online iNMF result for current datasets. "
                             "Please run online_inmf without newDatasets first")
    else:
        prev, data = [], list(objectList)
        if (Winit is None) != (Vinit is None):
            raise ValueError("Winit and Vinit must both be given, or both be None")
    npv, d = len(prev), len(prev) + len(data)
    if chunk_source is None:
        cs = [_Csc(m) for m in data]
        structs = [c.struct for c in cs]
        shapes = [c.shape for c in cs]
    else:
        shapes = [tuple(int(x) for x in sh) for sh in chunk_source["shapes"]]
        if len(shapes) != len(data):
            raise ValueError("chunk_source['shapes'] needs one (m, n_i) per dataset that is read")
        structs = []
        for sh in shapes:
            st = lgr_csc()
            st.nrow, st.ncol = sh
            structs.append(st)
    m = shapes[0][0]
    ns = [0] * npv + [sh[1] for sh in shapes]
    if min(ns[npv:]) < k:
        raise ValueError(f"Number of factors (k={k}) should be less than the number of cells in the smallest dataset ({min(ns[npv:])}).")
    if m < k:
        raise ValueError(f"Number of factors (k={k}) should be less than the number of shared features ({m}).")
    Earr = (lgr_csc * d)(*([lgr_csc() for _ in range(npv)] + structs))
    a = _ffi.lgr_online_args()
    a.d, a.k, a.m, a.E, a.lambda_ = d, int(k), m, C.cast(Earr, C.c_void_p), float(lambda_)
    a.max_epochs, a.minibatch_size, a.hals_iter, a.seed, a.device = int(maxEpochs), int(minibatchSize), int(HALSiter), int(seed), device
    a.n_prev, a.flags = npv, (1 if projection else 0)
    keep = []

    def ptr_array(lst):   # d pointers; None entries stay NULL
        vals = [None if x is None else _f64(x) for x in list(lst) + [None] * (d - len(lst))]
        arr = (c_double_p * d)(*[C.cast(None, c_double_p) if v is None else _dp(v) for v in vals])
        keep.extend([vals, arr])
        return arr
    if Winit is not None:
        W0 = _f64(Winit)
        keep.append(W0)
        a.Winit = _dp(W0)
    if Vinit is not None and not projection:
        a.Vinit = ptr_array(Vinit)
    if npv:
        a.Ainit = ptr_array(Ainit)
        a.Binit = ptr_array(Binit)
    W = np.zeros((m, k), order="F")
    V = [np.zeros((m, k), order="F") for _ in range(d)]
    H = [np.zeros((n, k), order="F") for n in ns]
    A = [np.zeros((k, k), order="F") for _ in range(d)]
    B = [np.zeros((m, k), order="F") for _ in range(d)]
    o = _ffi.lgr_online_out()
    arrs = [(c_double_p * d)(*[_dp(x) for x in lst]) for lst in (V, H, A, B)]
    o.W, o.V, o.H, o.A, o.B = _dp(W), arrs[0], arrs[1], arrs[2], arrs[3]
    if chunk_source is None:
        _ffi.check(_ffi.lib().lgr_online_inmf(C.byref(a), C.byref(o)))
    else:
        read = chunk_source["read"]
        failure = []

        def _cb(user, ds, c0, nc, colptr, rowidx, values, cap):   # called on this thread, inside lgr_online_inmf_chunked
            try:
                cp, ri, vv = read(int(ds) - npv, int(c0), int(nc))
                cp = np.asarray(cp, dtype=np.int64)
                got = int(cp[-1] - cp[0])
                if got > cap:
                    return -got
                np.ctypeslib.as_array(colptr, (int(nc) + 1,))[:] = cp - cp[0]
                if got:
                    np.ctypeslib.as_array(rowidx, (got,))[:] = np.asarray(ri, dtype=np.int32)[:got]
                    np.ctypeslib.as_array(values, (got,))[:] = np.asarray(vv, dtype=np.float64)[:got]
                return got
            except Exception as e:      # an exception must not cross the C frames
                failure.append(e)
                return -1
        src = _ffi.lgr_chunk_source()
        cb = _ffi.READ_COLS_FN(_cb)
        nnz_arr = (C.c_int64 * d)(*([0] * npv + [int(x) for x in chunk_source["nnz"]]))
        src.read_cols, src.user, src.nnz = cb, None, nnz_arr
        src.chunk_cols = int(chunk_source.get("chunk_cols", 1000))
        src.max_chunk_nnz = int(chunk_source.get("max_chunk_nnz", 0))
        code = _ffi.lib().lgr_online_inmf_chunked(C.byref(a), C.byref(src), C.byref(o))
        if failure:
            raise failure[0]
        _ffi.check(code)
    if projection:
        return {"H": H, "objErr": float(o.objErr), "loop_ms": float(o.loop_ms)}
    for i in range(npv):   # the already factorised datasets keep the H they came with
        H[i] = None if Hinit is None or Hinit[i] is None else np.asarray(Hinit[i])
    return {"H": H, "V": V, "W": W, "A": A, "B": B, "objErr": float(o.objErr), "iterations": int(o.iterations),
            "loop_ms": float(o.loop_ms)}


def bppnnls_prod(CtC, CtB, nCores=2, device=-1):
    """Drop-in for RcppPlanc::bppnnls_prod(CtC, CtB, nCores) (R/cINMF.R:481): returns X (k x c) >= 0."""
    G = _f64(CtC)
    B = _f64(CtB)
    if B.ndim == 1:
        B = _f64(B[:, None])
    k, c = B.shape
    if G.shape != (k, k):
        raise ValueError("CtC must be k x k and CtB k x c")
    X = np.zeros((k, c), order="F")
    _ffi.check(_ffi.lib().lgr_bppnnls_prod(k, c, _dp(G), _dp(B), _dp(X), device))
    return X


def objErr_i(H, W, V, E, lambda_, device=-1):
    """Drop-in for objErr_i (src/objErr.cpp:13-31): H is k x n, W/V are m x k, E is m x n sparse."""
    H = _f64(H)   # k x n column-major == [cell][k]
    W = _f64(W)
    V = _f64(V)
    e = _Csc(E)
    k = H.shape[0]
    if H.shape[1] != e.shape[1] or W.shape != (e.shape[0], k) or V.shape != W.shape:
        raise ValueError("objErr_i: inconsistent shapes")
    out = C.c_double(0.0)
    _ffi.check(_ffi.lib().lgr_objerr_i(k, _dp(H), _dp(W), _dp(V), C.byref(e.struct), float(lambda_), device,
                                       C.byref(out)))
    return float(out.value)


class Context:
    """Resident-handle API: inputs stay in HBM between calls (bench.py, multi-GPU hosts)."""

    def __init__(self, objectList, k, lambda_, Winit, Vinit, unsharedList=None, Uinit=None, Hinit=None, device=-1,
                 rank=0, world=1, nccl_id=None, flags=0, keep_f32=False):
        self.args = _Args(objectList, k, lambda_, 0, Winit, Vinit, Hinit, unsharedList=unsharedList, Uinit=Uinit,
                          flags=flags, device=device, rank=rank, world=world, nccl_id=nccl_id, keep_f32=keep_f32)
        self.h = C.c_void_p()
        _ffi.check(_ffi.lib().lgr_ctx_create(C.byref(self.args.struct), C.byref(self.h)))

    def reset(self, Winit, Vinit, Uinit=None):
        W0 = _f64(Winit)
        V0 = [_f64(v) for v in Vinit]
        d = self.args.d
        Varr = (c_double_p * d)(*[_dp(v) for v in V0])
        Uarr = None
        if Uinit is not None:
            U0 = [None if u is None else _f64(u) for u in Uinit]
            Uarr = (c_double_p * d)(*[(_dp(x) if x is not None else None) for x in U0])
        _ffi.check(_ffi.lib().lgr_ctx_reset(self.h, _dp(W0), Varr, Uarr))

    def iterate(self, niter, trajectory=False):
        traj = np.zeros(max(niter, 1)) if trajectory else None
        _ffi.check(_ffi.lib().lgr_ctx_iterate(self.h, int(niter), _dp(traj) if trajectory else None))
        return traj[:niter] if trajectory else None

    def objective(self):
        out = C.c_double(0.0)
        _ffi.check(_ffi.lib().lgr_ctx_objective(self.h, C.byref(out)))
        return float(out.value)

    def solve(self, blocks):
        """Block solves on the resident data (inmf_solveH / inmf_solveV / inmf_solveW, R/cINMF.R:477-503): `blocks` is a
        string over "HVW" (run in the order H -> V -> W), each from the factors currently on the device."""
        bits = 0
        for ch in blocks.upper():
            bits |= {"H": _ffi.SOLVE_H, "V": _ffi.SOLVE_V, "W": _ffi.SOLVE_W}[ch]
        _ffi.check(_ffi.lib().lgr_ctx_solve(self.h, bits))

    def set_factors(self, W=None, V=None, U=None, H=None):
        """Replace resident factors (None keeps): W m x k, V[i] m x k, U[i] u_i x k, H[i] n_i x k (cell x factor)."""
        d = self.args.d
        keep = []

        def arr(lst):
            if lst is None:
                return None
            mats = [None if x is None else _f64(x) for x in lst]
            keep.append(mats)
            return (c_double_p * d)(*[(_dp(x) if x is not None else None) for x in mats])
        Wm = None if W is None else _f64(W)
        _ffi.check(_ffi.lib().lgr_ctx_set_factors(self.h, _dp(Wm) if Wm is not None else None, arr(V), arr(U), arr(H)))

    def products(self, which):
        """which = "ltx": [E_i; P_i]^T ([W; 0] + [V_i; U_i]) per dataset (n_i x k); which = "xh": [E_i; P_i] H_i
        ((m + u_i) x k) -- the two sweeps over the resident X, for callers that assemble their own normal equations."""
        d, k = self.args.d, self.args.k
        if which == "ltx":
            outs = [np.zeros((n, k), order="F") for n in self.args.n]
        elif which == "xh":
            outs = [np.zeros((self.args.m + u, k), order="F") for u in self.args.u]
        else:
            raise ValueError("which must be 'ltx' or 'xh'")
        arr = (c_double_p * d)(*[_dp(o) for o in outs])
        _ffi.check(_ffi.lib().lgr_ctx_products(self.h, 0 if which == "ltx" else 1, arr))
        return outs

    def objErr_i(self, i):
        """objErr_i (src/objErr.cpp:13-31) of dataset i from the resident factors."""
        out = C.c_double(0.0)
        _ffi.check(_ffi.lib().lgr_ctx_objerr_i(self.h, int(i), C.byref(out)))
        return float(out.value)

    def run(self, niter):
        """niter iterations + the final objective inside one device-timed bracket (what lgr_inmf runs between its
        upload and its download); returns the objective."""
        out = C.c_double(0.0)
        _ffi.check(_ffi.lib().lgr_ctx_run(self.h, int(niter), C.byref(out)))
        return float(out.value)

    def download(self, want_passive=False):
        o = _Out(self.args, want_passive=want_passive)
        _ffi.check(_ffi.lib().lgr_ctx_download(self.h, C.byref(o.struct)))
        res = {"H": o.H, "V": o.V, "W": o.W, "U": o.U, "timings": o.timings()}
        if want_passive:
            res["H_passive"] = o.passive
        return res

    def timings(self):
        t = _ffi.lgr_timings()
        _ffi.check(_ffi.lib().lgr_ctx_timings(self.h, C.byref(t)))
        return {f: getattr(t, f) for f, _ in _ffi.lgr_timings._fields_}

    def set_kernel_timing(self, on=True):
        _ffi.check(_ffi.lib().lgr_ctx_set_kernel_timing(self.h, 1 if on else 0))

    def kernel_times(self):
        n = C.c_int32(32)
        names = (C.c_char_p * 32)()
        ms = (C.c_double * 32)()
        cnt = (C.c_int64 * 32)()
        _ffi.check(_ffi.lib().lgr_ctx_kernel_times(self.h, C.byref(n), names, ms, cnt))
        return {names[i].decode(): {"ms": ms[i], "launches": cnt[i]} for i in range(min(n.value, 32))}

    def stream(self):
        return _ffi.lib().lgr_ctx_stream(self.h)

    def close(self):
        if self.h:
            _ffi.lib().lgr_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    _ffi.check(_ffi.lib().lgr_nccl_unique_id(buf))
    return buf.raw


def device_count():
    return int(_ffi.lib().lgr_device_count())


class RRandom:
    """R's `set.seed(seed); runif(n, lo, hi)` stream (host C++ in csrc/lgr_rng.cpp)."""

    def __init__(self, seed):
        self._h = _ffi.lib().lgr_r_set_seed(int(seed) & 0xFFFFFFFF)

    def runif(self, n, lo=0.0, hi=1.0):
        out = np.empty(int(n), dtype=np.float64)
        _ffi.lib().lgr_r_runif(self._h, int(n), float(lo), float(hi), _dp(out))
        return out

    def __del__(self):
        try:
            _ffi.lib().lgr_r_free(self._h)
        except Exception:
            pass
