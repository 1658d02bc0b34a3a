"""Mirror of the preprocessing that produces the hot path's input, `scaleData`.

    normalize       <- normalize.dgCMatrix                  (reference R/preprocess.R:592-608)
    selectGenes     <- .selectGenes.ligerDataset + .selectGenes.withMetric, default arguments
                                                            (R/preprocess.R:1005-1057, 1184-1213)
    scaleNotCenter  <- scaleNotCenter.dgCMatrix             (R/preprocess.R:1456-1472)

The sweeps over the matrices -- column sums and division, per-gene mean / variance (rowVars_sparse_rcpp,
src/data_processing.cpp:152-172), per-gene root-mean-square scaling (scaleNotCenter_byRow_rcpp, :42-58) -- run on the GPU
through the C ABI (lgr_normalize_csc, lgr_row_vars_sparse, lgr_scale_not_center_by_row).  The per-gene threshold test and
the set logic over gene names are host code, as they are R code in the reference.
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp
from scipy.stats import norm as _norm

from . import _ffi
from ._ffi import c_double_p, c_int32_p


def _csc(x):
    x = sp.csc_matrix(x, dtype=np.float64)
    x.sort_indices()
    return x


def _ip(a):
    return a.ctypes.data_as(c_int32_p)


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def normalize(raw, device=-1):
    """Each column divided by its sum (library-size normalisation).  Returns a CSC matrix with the same pattern."""
    raw = _csc(raw)
    indptr = np.ascontiguousarray(raw.indptr, dtype=np.int32)
    data = np.ascontiguousarray(raw.data, dtype=np.float64)
    out = np.empty_like(data)
    _ffi.check(_ffi.lib().lgr_normalize_csc(raw.shape[1], _ip(indptr), _dp(data), _dp(out), None, device))
    return sp.csc_matrix((out, raw.indices.copy(), raw.indptr.copy()), shape=raw.shape)


def rowMeansVars(x, ncol=None, means=None, device=-1):
    """(Matrix::rowMeans(x), rowVars_sparse_rcpp(x, means, ncol)) of a sparse genes x cells matrix."""
    x = _csc(x)
    nrow, ncx = x.shape
    indptr = np.ascontiguousarray(x.indptr, dtype=np.int32)
    indices = np.ascontiguousarray(x.indices, dtype=np.int32)
    data = np.ascontiguousarray(x.data, dtype=np.float64)
    m_out = np.empty(nrow)
    v_out = np.empty(nrow)
    m_in = None if means is None else np.ascontiguousarray(means, dtype=np.float64)
    _ffi.check(_ffi.lib().lgr_row_vars_sparse(nrow, ncx, _ip(indptr), _ip(indices), _dp(data),
                                              None if m_in is None else _dp(m_in), float(ncx if ncol is None else ncol),
                                              _dp(m_out), _dp(v_out), device))
    return m_out, v_out


def selectGenesOne(norm, genes, sharedMask, nUMI, thresh=0.1, alpha=0.99, device=-1):
    """Variable genes of one dataset among the shared genes (R/preprocess.R:1024-1057, 1184-1213)."""
    means, vars_ = rowMeansVars(norm, device=device)
    sharedMask = np.asarray(sharedMask, dtype=bool)
    means, vars_ = means[sharedMask], vars_[sharedMask]
    g = [x for x, s in zip(genes, sharedMask) if s]
    nolan = float(np.mean(1.0 / np.asarray(nUMI, dtype=np.float64)))
    a_corr = alpha / norm.shape[0]
    upper = means + _norm.ppf(1.0 - a_corr / 2.0) * np.sqrt(means * nolan / norm.shape[1])
    with np.errstate(divide="ignore", invalid="ignore"):
        base = np.log10(means * nolan)
        keep = (vars_ / nolan > upper) & (np.log10(vars_) > base + thresh)
    return [x for x, s in zip(g, keep) if s]


def selectGenes(norms, genesList, nUMIs, thresh=0.1, alpha=0.99, combine="union", device=-1):
    """selectGenes over a list of normalised matrices: per-dataset selection among the genes shared by all datasets,
    then union (in order of first appearance) or intersection (R/preprocess.R:1040-1057)."""
    shared = set(genesList[0])
    for g in genesList[1:]:
        shared &= set(g)
    sel = [selectGenesOne(n, g, [x in shared for x in g], u, thresh, alpha, device)
           for n, g, u in zip(norms, genesList, nUMIs)]
    if combine == "union":
        feats, seen = [], set()
        for s in sel:
            for x in s:
                if x not in seen:
                    seen.add(x)
                    feats.append(x)
    elif combine == "intersection":
        feats = [x for x in sel[0] if all(x in set(s) for s in sel[1:])]
    else:
        raise ValueError("`combine` must be 'union' or 'intersection'")
    return feats, sel


def scaleNotCenter(norm, genes, features, device=-1):
    """Rows `features` of the normalised matrix, each divided by its root mean square over ALL cells (not centred).
    Returns the features x cells CSC `scaleData` (explicit zeros dropped)."""
    pos = {g: i for i, g in enumerate(genes)}
    try:
        rows = np.array([pos[f] for f in features], dtype=np.int64)
    except KeyError as e:
        raise ValueError(f"feature {e.args[0]!r} not found in the dataset") from None
    sub = _csc(sp.csr_matrix(norm)[rows, :])
    if sub.shape[0] == 0:
        return sub
    indptr = np.ascontiguousarray(sub.indptr, dtype=np.int32)
    indices = np.ascontiguousarray(sub.indices, dtype=np.int32)
    data = np.ascontiguousarray(sub.data, dtype=np.float64)
    out = np.empty_like(data)
    _ffi.check(_ffi.lib().lgr_scale_not_center_by_row(sub.shape[0], sub.shape[1], _ip(indptr), _ip(indices), _dp(data),
                                                      _dp(out), None, device))
    out[np.isnan(out)] = 0.0                    # scaled@x[is.na(scaled@x)] <- 0   (R/preprocess.R:1469)
    res = sp.csc_matrix((out, sub.indices.copy(), sub.indptr.copy()), shape=sub.shape)
    res.eliminate_zeros()
    return res


class Pipeline:
    """normalize -> selectGenes -> scaleNotCenter with the matrices RESIDENT in HBM (lgr_prep_*): the raw counts are
    uploaded once, only the per-gene statistics travel back to the host (the variable-gene rule and the set logic over
    gene names are host code, as they are R code in the reference), and the resulting scaleData is handed to
    inmf / uinmf / Context as device views -- X is never copied back nor uploaded a second time.

        pipe = Pipeline(raws, genesList, nUMIs)
        features = pipe.selectGenes()            # normalises first if needed
        Es = pipe.scaleNotCenter(features)       # list of api.DeviceCsc
        fit = liger_b200.run_inmf_list(Es, k=20, ...)
    """

    def __init__(self, raws, genesList, nUMIs, device=-1):
        from . import api
        self._api = api
        self.genes = [list(g) for g in genesList]
        self.nUMI = [np.asarray(u, dtype=np.float64) for u in nUMIs]
        # scipy CSC matrices or (indptr, indices, data, shape) tuples (e.g. page-locked buffers): used in place, no copy
        self._host = [api._Csc(r) for r in raws]
        arr = (_ffi.lgr_csc * len(raws))(*[h.struct for h in self._host])
        self.h = C.c_void_p()
        _ffi.check(_ffi.lib().lgr_prep_create(len(raws), arr, device, C.byref(self.h)))
        self.shapes = [h.shape for h in self._host]
        self._host = None
        self._normalized = False

    def normalize(self):
        if not self._normalized:
            _ffi.check(_ffi.lib().lgr_prep_normalize(self.h))
            self._normalized = True
        return self

    def geneStats(self, i):
        self.normalize()
        m = np.empty(self.shapes[i][0])
        v = np.empty(self.shapes[i][0])
        _ffi.check(_ffi.lib().lgr_prep_gene_stats(self.h, i, _dp(m), _dp(v)))
        return m, v

    def selectGenes(self, thresh=0.1, alpha=0.99, combine="union"):
        """Same rule as selectGenes() above; returns (features, per-dataset selections)."""
        shared = set(self.genes[0])
        for g in self.genes[1:]:
            shared &= set(g)
        sel = []
        for i, g in enumerate(self.genes):
            means, vars_ = self.geneStats(i)
            mask = np.array([x in shared for x in g], dtype=bool)
            means, vars_ = means[mask], vars_[mask]
            gs = [x for x, s_ in zip(g, mask) if s_]
            ncol = self.shapes[i][1]
            nolan = float(np.mean(1.0 / self.nUMI[i]))
            a_corr = alpha / self.shapes[i][0]
            upper = means + _norm.ppf(1.0 - a_corr / 2.0) * np.sqrt(means * nolan / ncol)
            with np.errstate(divide="ignore", invalid="ignore"):
                keep = (vars_ / nolan > upper) & (np.log10(vars_) > np.log10(means * nolan) + thresh)
            sel.append([x for x, s_ in zip(gs, keep) if s_])
        if combine == "union":
            feats, seen = [], set()
            for s_ in sel:
                for x in s_:
                    if x not in seen:
                        seen.add(x)
                        feats.append(x)
        elif combine == "intersection":
            feats = [x for x in sel[0] if all(x in set(s_) for s_ in sel[1:])]
        else:
            raise ValueError("`combine` must be 'union' or 'intersection'")
        return feats, sel

    def scaleNotCenter(self, features):
        """Builds scaleData of every dataset on the device; returns the list of device views."""
        self.normalize()
        views = []
        for i, g in enumerate(self.genes):
            pos = {x: j for j, x in enumerate(g)}
            try:
                rows = np.array([pos[f] for f in features], dtype=np.int32)
            except KeyError as e:
                raise ValueError(f"feature {e.args[0]!r} not found in dataset {i}") from None
            _ffi.check(_ffi.lib().lgr_prep_scale(self.h, i, len(rows), _ip(rows)))
            st = _ffi.lgr_csc()
            _ffi.check(_ffi.lib().lgr_prep_scaled_view(self.h, i, C.byref(st)))
            views.append(self._api.DeviceCsc(st, self))
        self._views = views
        return views

    def scaledToHost(self, i):
        """Host copy of scaleData_i (FP64 values) -- for checks; the factorisation does not need it."""
        v = self._views[i].struct
        n = int(v.ncol)
        indptr = np.empty(n + 1, dtype=np.int32)
        _ffi.check(_ffi.lib().lgr_prep_scaled_download(self.h, i, _ip(indptr), None, None))
        nnz = int(indptr[-1])
        indices = np.empty(max(nnz, 1), dtype=np.int32)
        data = np.empty(max(nnz, 1), dtype=np.float64)
        _ffi.check(_ffi.lib().lgr_prep_scaled_download(self.h, i, None, _ip(indices), _dp(data)))
        out = sp.csc_matrix((data[:nnz], indices[:nnz], indptr), shape=(int(v.nrow), n))
        out.eliminate_zeros()
        out.sort_indices()
        return out

    def close(self):
        if getattr(self, "h", None):
            _ffi.lib().lgr_prep_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
