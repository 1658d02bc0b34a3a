"""Mirror of the reference's quantileNorm (R/integration.R:1470-1711): joint clustering by arg-max factor, kNN refinement
and the per-cluster, per-factor quantile warp of every dataset onto the reference dataset.  The whole of
`.quantileNorm.HList` runs in the library (csrc/lgr_align.cu: lgr_quantile_norm, or lgr_ctx_quantile_norm on an H that is
still resident after the factorisation); this file only holds the argument handling of the R front ends.

    quantileNorm(fit_or_HList, quantiles=50, reference=None, minCells=20, nNeighbors=20, useDims=None, center=False,
                 maxSample=1000, eps=0.9, refineKNN=True, clusterName="quantileNorm_cluster", seed=1)

`fit` is what integration.runINMF returns (H = {name: k x n_i}); an HList is {name: k x n_i} (or a list).  Returns the fit
with `H.norm` (sum n_i x k, datasets stacked in order), `clusters` ({name: labels}) and uns["alignmentMethod"] added, or
for an HList dict('H.norm', 'clusters') like the internal R function.
"""
import ctypes as C

import numpy as np

from . import _ffi
from ._ffi import c_double_p


def _resolve_reference(reference, names):
    # R/integration.R:1636-1649
    if isinstance(reference, str):
        if reference not in names:
            raise ValueError("Should specify one existing dataset as reference.")
        return names.index(reference)
    if isinstance(reference, (bool, np.bool_)) or (isinstance(reference, (list, tuple, np.ndarray)) and
                                                   len(reference) and isinstance(reference[0], (bool, np.bool_))):
        ref = np.atleast_1d(np.asarray(reference, dtype=bool))
        if ref.shape[0] != len(names) or ref.sum() != 1:
            raise ValueError("Should specify one existing dataset as reference.")
        return int(np.argmax(ref))
    if isinstance(reference, (int, np.integer, float)):
        if reference > len(names) or reference < 1:
            raise ValueError("Should specify one existing dataset as reference.")
        return int(reference) - 1          # R's 1-based index
    raise ValueError("Unable to understand `reference`. See `?quantileNorm`.")


def _params(k, reference, quantiles, minCells, nNeighbors, useDims, center, maxSample, eps, refineKNN, seed):
    p = _ffi.lgr_qn_params()
    p.quantiles, p.reference, p.min_cells, p.n_neighbors = int(quantiles), int(reference), int(minCells), int(nNeighbors)
    p.center, p.max_sample, p.refine_knn, p.seed, p.eps = int(bool(center)), int(maxSample), int(bool(refineKNN)), int(seed), float(eps)
    keep = None
    if useDims is not None:
        keep = np.ascontiguousarray(useDims, dtype=np.int32)
        p.n_dims_use = keep.shape[0]
        p.dims_use = keep.ctypes.data_as(C.POINTER(C.c_int32))
    return p, keep


def quantile_norm_hlist(HList, quantiles=50, reference=None, minCells=20, nNeighbors=20, useDims=None, center=False,
                        maxSample=1000, eps=0.9, refineKNN=True, seed=1, device=-1, return_info=False):
    """.quantileNorm.HList: HList = {name: k x n_i} (or a list of k x n_i)."""
    names = list(HList.keys()) if isinstance(HList, dict) else [str(i + 1) for i in range(len(HList))]
    mats = list(HList.values()) if isinstance(HList, dict) else list(HList)
    if reference is None:
        raise ValueError("Should specify one existing dataset as reference.")
    ref = _resolve_reference(reference, names)
    k = mats[0].shape[0]
    Hs = [np.asfortranarray(np.asarray(h, dtype=np.float64).T) for h in mats]      # cell x factor
    for h in Hs:
        if h.shape[1] != k:
            raise ValueError("All H matrices must have the same number of factors.")
    d = len(Hs)
    p, _keep = _params(k, ref, quantiles, minCells, nNeighbors, useDims, center, maxSample, eps, refineKNN, seed)
    outs = [np.zeros(h.shape, order="F") for h in Hs]
    labs = [np.zeros(h.shape[0], dtype=np.int32) for h in Hs]
    n = (C.c_int64 * d)(*[h.shape[0] for h in Hs])
    Harr = (c_double_p * d)(*[h.ctypes.data_as(c_double_p) for h in Hs])
    Oarr = (c_double_p * d)(*[o.ctypes.data_as(c_double_p) for o in outs])
    Larr = (C.POINTER(C.c_int32) * d)(*[x.ctypes.data_as(C.POINTER(C.c_int32)) for x in labs])
    info = _ffi.lgr_qn_info()
    _ffi.check(_ffi.lib().lgr_quantile_norm(d, k, n, Harr, C.byref(p), Oarr, Larr, C.byref(info), device))
    res = {"H.norm": np.vstack(outs), "clusters": dict(zip(names, labs))}
    if return_info:
        res["info"] = {f: getattr(info, f) for f, _ in _ffi.lgr_qn_info._fields_ if not f.startswith("pad")}
    return res


def quantile_norm_resident(ctx, reference=0, quantiles=50, minCells=20, nNeighbors=20, useDims=None, center=False,
                           maxSample=1000, eps=0.9, refineKNN=True, seed=1):
    """quantileNorm of the H_i that are still in HBM after the factorisation (api.Context); reference is a 0-based index."""
    d, k = ctx.args.d, ctx.args.k
    p, _keep = _params(k, reference, quantiles, minCells, nNeighbors, useDims, center, maxSample, eps, refineKNN, seed)
    outs = [np.zeros((n, k), order="F") for n in ctx.args.n]
    labs = [np.zeros(n, dtype=np.int32) for n in ctx.args.n]
    Oarr = (c_double_p * d)(*[o.ctypes.data_as(c_double_p) for o in outs])
    Larr = (C.POINTER(C.c_int32) * d)(*[x.ctypes.data_as(C.POINTER(C.c_int32)) for x in labs])
    info = _ffi.lgr_qn_info()
    _ffi.check(_ffi.lib().lgr_ctx_quantile_norm(ctx.h, C.byref(p), Oarr, Larr, C.byref(info)))
    return {"H.norm": np.vstack(outs), "clusters": labs,
            "info": {f: getattr(info, f) for f, _ in _ffi.lgr_qn_info._fields_ if not f.startswith("pad")}}


def quantileNorm(obj, quantiles=50, reference=None, minCells=20, nNeighbors=20, useDims=None, center=False, maxSample=1000,
                 eps=0.9, refineKNN=True, clusterName="quantileNorm_cluster", seed=1, verbose=False):
    """quantileNorm.liger (R/integration.R:1520-1565) on a fit dict, or the list method on an HList."""
    is_fit = isinstance(obj, dict) and "H" in obj and "W" in obj
    HList = obj["H"] if is_fit else obj
    names = list(HList.keys()) if isinstance(HList, dict) else None
    if reference is None:
        # .autoFindRef_qn (R/integration.R:1713-1730) without modality classes: the dataset with the most cells
        sizes = [np.asarray(h).shape[1] for h in (HList.values() if names else HList)]
        reference = int(np.argmax(sizes)) + 1
    res = quantile_norm_hlist(HList, quantiles=quantiles, reference=reference, minCells=minCells, nNeighbors=nNeighbors,
                              useDims=useDims, center=center, maxSample=maxSample, eps=eps, refineKNN=refineKNN, seed=seed)
    if not is_fit:
        return res
    out = dict(obj)
    out["H.norm"] = res["H.norm"]
    out.setdefault("cellMeta", {})
    out["cellMeta"] = dict(out["cellMeta"])
    out["cellMeta"][clusterName] = res["clusters"]
    out["uns"] = dict(out.get("uns", {}))
    out["uns"]["alignmentMethod"] = "quantileNorm"
    return out


# ------------------------------------------------------------------------------------------------------------------
# centroidAlign (R/integration.R:1880-2084)
# ------------------------------------------------------------------------------------------------------------------
def _ca_params(d, lambda_, useDims, scaleEmb, centerEmb, scaleCluster, centerCluster, shift):
    lam = np.ascontiguousarray(np.full(d, float(lambda_)) if np.ndim(lambda_) == 0 else np.asarray(lambda_, dtype=np.float64))
    if lam.shape[0] != d:      # .checkArgLen(lambda, length(object), repN = TRUE)
        raise ValueError(f"`lambda` has to be a length 1 or {d} numeric vector.")
    p = _ffi.lgr_ca_params()
    p.lambda_ = lam.ctypes.data_as(c_double_p)
    p.scale_emb, p.center_emb, p.scale_cluster, p.center_cluster = int(bool(scaleEmb)), int(bool(centerEmb)), int(bool(scaleCluster)), int(bool(centerCluster))
    p.shift = int(bool(shift))
    keep = [lam]
    if useDims is not None:
        dims = np.ascontiguousarray(useDims, dtype=np.int32)
        p.n_dims_use = dims.shape[0]
        p.dims_use = dims.ctypes.data_as(C.POINTER(C.c_int32))
        keep.append(dims)
    return p, keep


def centroid_align_hlist(HList, lambda_=1.0, useDims=None, scaleEmb=True, centerEmb=True, scaleCluster=False,
                         centerCluster=False, shift=False, diagnosis=False, device=-1):
    """.centroidAlign.Hraw on {name: k x n_i} (datasets in order = the levels of datasetVar)."""
    mats = list(HList.values()) if isinstance(HList, dict) else list(HList)
    k = mats[0].shape[0]
    Hs = [np.asfortranarray(np.asarray(h, dtype=np.float64).T) for h in mats]
    d = len(Hs)
    p, _keep = _ca_params(d, lambda_, useDims, scaleEmb, centerEmb, scaleCluster, centerCluster, shift)
    kd = k if useDims is None else len(useDims)
    outs = [np.zeros((h.shape[0], kd), order="F") for h in Hs]
    n = (C.c_int64 * d)(*[h.shape[0] for h in Hs])
    Harr = (c_double_p * d)(*[h.ctypes.data_as(c_double_p) for h in Hs])
    Oarr = (c_double_p * d)(*[o.ctypes.data_as(c_double_p) for o in outs])
    ip = C.POINTER(C.c_int32)
    diag = [[np.zeros(h.shape[0], dtype=np.int32) for h in Hs] for _ in range(3)] if diagnosis else None
    darr = [((ip * d)(*[x.ctypes.data_as(ip) for x in g]) if diagnosis else None) for g in (diag or [None] * 3)]
    _ffi.check(_ffi.lib().lgr_centroid_align(d, k, n, Harr, C.byref(p), Oarr, darr[0], darr[1], darr[2], device))
    res = {"aligned": np.vstack(outs)}
    if diagnosis:
        res["raw_which.max"], res["Z_which.max"], res["R_which.max"] = (np.concatenate(g) for g in diag)
    return res


def centroid_align_resident(ctx, lambda_=1.0, useDims=None, scaleEmb=True, centerEmb=True, scaleCluster=False,
                            centerCluster=False, shift=False):
    d, k = ctx.args.d, ctx.args.k
    p, _keep = _ca_params(d, lambda_, useDims, scaleEmb, centerEmb, scaleCluster, centerCluster, shift)
    kd = k if useDims is None else len(useDims)
    outs = [np.zeros((n, kd), order="F") for n in ctx.args.n]
    Oarr = (c_double_p * d)(*[o.ctypes.data_as(c_double_p) for o in outs])
    _ffi.check(_ffi.lib().lgr_ctx_centroid_align(ctx.h, C.byref(p), Oarr))
    return {"aligned": np.vstack(outs)}


def centroidAlign(obj, lambda_=1.0, useDims=None, scaleEmb=True, centerEmb=True, scaleCluster=False, centerCluster=False,
                  shift=False, diagnosis=False):
    """centroidAlign.liger (R/integration.R:1898-1950) on a fit dict (H = {name: k x n_i}), or on an HList."""
    is_fit = isinstance(obj, dict) and "H" in obj and "W" in obj
    res = centroid_align_hlist(obj["H"] if is_fit else obj, lambda_=lambda_, useDims=useDims, scaleEmb=scaleEmb,
                               centerEmb=centerEmb, scaleCluster=scaleCluster, centerCluster=centerCluster, shift=shift,
                               diagnosis=diagnosis)
    if not is_fit:
        return res
    out = dict(obj)
    out["H.norm"] = res["aligned"]
    out["uns"] = dict(out.get("uns", {}))
    out["uns"]["alignmentMethod"] = "centroidAlign"
    if diagnosis:
        out["cellMeta"] = dict(out.get("cellMeta", {}))
        for key in ("raw_which.max", "Z_which.max", "R_which.max"):
            out["cellMeta"][key] = res[key]
    return out


def alignFactors(obj, method="quantileNorm", **kw):
    """alignFactors (R/integration.R:1400-1426): dispatch on the method name."""
    m = method.lower()
    if m in ("quantilenorm", "quantile_norm", "quantile"):
        return quantileNorm(obj, **kw)
    if m in ("centroidalign", "centroid_align", "centroid"):
        return centroidAlign(obj, **kw)
    raise ValueError("`method` must be one of 'quantileNorm', 'quantile_norm', 'centroidAlign', 'centroid_align'")
