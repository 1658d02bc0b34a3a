"""Mirror of the reference's "re-optimise from an existing factorisation" callers of the hot path
(reference R/optimizeNewParam.R).  They are host-level algebra around the two native entry points this library
replaces -- RcppPlanc::bppnnls_prod (-> lgr_bppnnls_prod) and RcppPlanc::inmf with explicit inits (-> lgr_inmf) --
so every NNLS solve and every ANLS iteration below runs on the GPU, and so do the products over the sparse data
(lgr_ctx_products / lgr_ctx_solve on a resident copy of E); only the k x k corrections stay on the host, as they do in R.

    optimizeNewK       R/optimizeNewParam.R:41-163
    optimizeNewLambda  R/optimizeNewParam.R:353-385
    optimizeNewData    R/optimizeNewParam.R:245-351   (on SCALED data: normalize / scaleNotCenter are outside the path)
    optimizeSubset     R/optimizeNewParam.R:433-489

A "fit" is what integration.runINMF returns: dict(W = m x k, H = {name: k x n_i}, V = {name: m x k},
uns = {factorization: {k, lambda}}); `scaleData` is the {name: genes x cells sparse} dict it was fitted on.
"""
import numpy as np
import scipy.sparse as sp

from . import api, cinmf
from .integration import runINMF


def _check_fit(fit, scaleData):
    # .checkValidFactorResult (R/optimizeNewParam.R:55): W, and H / V of every dataset, must exist and agree in k
    if fit is None or fit.get("W") is None:
        raise ValueError("W matrix does not exist. Please run factorization first (e.g. runINMF()).")
    k = fit["W"].shape[1]
    for name in scaleData:
        if name not in fit["H"] or name not in fit["V"]:
            raise ValueError(f"H or V matrix of dataset '{name}' does not exist.")
        if fit["H"][name].shape[0] != k or fit["V"][name].shape[1] != k:
            raise ValueError(f"Number of factors does not match in dataset '{name}'.")
        if fit["H"][name].shape[1] != scaleData[name].shape[1]:
            raise ValueError(f"H of dataset '{name}' does not match its cells.")
    return k


def _dense(x):
    return np.asarray(x.todense()) if sp.issparse(x) else np.asarray(x)


def optimizeNewK(scaleData, fit, kNew, lambda_=None, nIteration=30, seed=1, nCores=2, verbose=False):
    """New number of factors from an existing fit.  kNew > k: the kNew - k extra factors are initialised by three NNLS
    solves against the current residual (H_new, V_new, W_new; R/optimizeNewParam.R:78-137); kNew < k: the kNew factors
    with the largest total loading norm are kept (:141-150).  Then runINMF from those inits (:152-162)."""
    k = _check_fit(fit, scaleData)
    lam = fit["uns"]["factorization"]["lambda"] if lambda_ is None else lambda_
    if kNew == k:
        return fit
    names = list(scaleData.keys())
    E = [scaleData[n] for n in names]
    W = np.asarray(fit["W"], dtype=np.float64)
    V = [np.asarray(fit["V"][n], dtype=np.float64) for n in names]
    H = [np.asarray(fit["H"][n], dtype=np.float64) for n in names]          # k x n_i
    g = W.shape[0]
    if kNew > k:
        kd = kNew - k
        rng = api.RRandom(seed)
        # t(matrix(runif(g * kd, 0, 2), kd, g)): entry (gene, j) is draw gene * kd + j
        W_new = rng.runif(g * kd, 0.0, 2.0).reshape((g, kd))
        V_new = [rng.runif(g * kd, 0.0, 2.0).reshape((g, kd)) for _ in names]
        # the two products over E (E^T L_new, E H_new^T) come from the library's sweeps over a resident copy of E (one
        # upload, kd factors wide); the k x k / k x kd corrections stay host algebra, as in R
        ctx = api.Context(E, kd, lam, W_new, V_new)
        try:
            EtL = ctx.products("ltx")                                              # n_i x kd each
            H_new = []
            for i in range(len(names)):
                L_new = W_new + V_new[i]
                L = W + V[i]
                CtC = L_new.T @ L_new + lam * (V_new[i].T @ V_new[i])
                CtB = -1.0 * ((L_new.T @ L) @ H[i] - EtL[i].T)
                H_new.append(api.bppnnls_prod(CtC, CtB, nCores=nCores))            # kd x n_i
            ctx.set_factors(H=[h.T for h in H_new])
            EH = ctx.products("xh")                                                # g x kd each
        finally:
            ctx.close()
        for i in range(len(names)):
            Hn = H_new[i]
            CtC = (Hn @ Hn.T) * (1.0 + lam)
            CtB = EH[i].T - (Hn @ H[i].T) @ (W + V[i]).T - (Hn @ Hn.T) @ W_new.T
            V_new[i] = api.bppnnls_prod(CtC, CtB, nCores=nCores).T             # g x kd
        CtC = np.zeros((kd, kd))
        CtB = np.zeros((kd, g))
        for i in range(len(names)):
            Hn = H_new[i]
            CtC += Hn @ Hn.T
            CtB += EH[i].T - (Hn @ H[i].T) @ (W + V[i]).T - (Hn @ Hn.T) @ V_new[i].T
        W_new = api.bppnnls_prod(CtC, CtB, nCores=nCores).T
        HInit = [np.vstack([H[i], H_new[i]]).T for i in range(len(names))]     # n_i x kNew
        VInit = [np.hstack([V[i], V_new[i]]) for i in range(len(names))]
        WInit = np.hstack([W, W_new])
    else:
        deltas = np.zeros(k)
        for i in range(len(names)):
            L = W + V[i]
            # norm(outer(L[, ki], H[ki, ]), "F") = |L[, ki]| * |H[ki, ]|
            deltas += np.linalg.norm(L, axis=0) * np.linalg.norm(H[i], axis=1)
        k_use = np.argsort(-deltas, kind="stable")[:kNew]                       # order(deltas, decreasing = TRUE)
        WInit = W[:, k_use]
        VInit = [v[:, k_use] for v in V]
        HInit = [h[k_use, :].T for h in H]
    return runINMF(scaleData, k=kNew, lambda_=lam, nIteration=nIteration, nRandomStarts=1, HInit=HInit, WInit=WInit,
                   VInit=VInit, seed=seed, nCores=nCores, verbose=verbose)


def optimizeNewLambda(scaleData, fit, lambdaNew, nIteration=30, seed=1, nCores=2, verbose=False, warn=None):
    """Re-optimise with a new lambda from the current H and W; V is re-drawn from the seed because the reference
    does not pass VInit (R/optimizeNewParam.R:375-384)."""
    k = _check_fit(fit, scaleData)
    if lambdaNew < fit["uns"]["factorization"]["lambda"] and warn is not None:
        warn("New lambda less than current lambda; new factorization may not be optimal. "
             "Re-optimization with runINMF recommended instead.")
    names = list(scaleData.keys())
    return runINMF(scaleData, k=k, lambda_=lambdaNew, nIteration=nIteration,
                   HInit=[np.asarray(fit["H"][n]).T for n in names], WInit=np.asarray(fit["W"]), seed=seed,
                   nCores=nCores, verbose=verbose)


def optimizeNewData(scaleData, fit, scaleDataNew, useDatasets, merge=True, newCells=None, lambda_=None, nIteration=30,
                    seed=1, nCores=2, verbose=False):
    """New cells from an existing fit (R/optimizeNewParam.R:245-351), on scaled data.

    merge=True : scaleDataNew[name] is the re-scaled MERGED matrix of dataset useDatasets[i] (old cells first) and
                 newCells[name] the column indices of the added cells; their H columns come from one NNLS solve against
                 the dataset's current W + V (:291-303) and are appended to the existing H.
    merge=False: scaleDataNew[name] is a new dataset; it inherits V from useDatasets[i] and its H from the same NNLS
                 solve over all its cells (:330-337).
    Then runINMF over all datasets from those inits (:341-350)."""
    k = _check_fit(fit, scaleData)
    lam = fit["uns"]["factorization"]["lambda"] if lambda_ is None else lambda_
    new_names = list(scaleDataNew.keys())
    if len(useDatasets) != len(new_names):
        raise ValueError("Length and order of `useDatasets` should match with `dataNew`.")
    for u in useDatasets:
        if u not in scaleData:
            raise ValueError(f"Dataset '{u}' not found.")
    W = np.asarray(fit["W"], dtype=np.float64)
    data = dict(scaleData)
    Hs = {n: np.asarray(fit["H"][n], dtype=np.float64) for n in scaleData}
    Vs = {n: np.asarray(fit["V"][n], dtype=np.float64) for n in scaleData}

    def h_for(E, V):                        # inmf_solveH on the new cells, k x n
        return cinmf.inmf_solveH(None, W, V, E, lam, nCores).T

    if merge:
        if newCells is None:
            raise ValueError("merge=True needs `newCells` (column indices of the added cells in each merged matrix)")
        for nm, tgt in zip(new_names, useDatasets):
            E = sp.csc_matrix(scaleDataNew[nm])
            idx = np.asarray(newCells[nm])
            Hs[tgt] = np.hstack([Hs[tgt], h_for(E[:, idx], Vs[tgt])])
            data[tgt] = E
    else:
        for nm, src in zip(new_names, useDatasets):
            if nm in scaleData:
                raise ValueError("Names of `dataNew` must be unique and different from existing datasets.")
            E = sp.csc_matrix(scaleDataNew[nm])
            data[nm] = E
            Vs[nm] = Vs[src].copy()
            Hs[nm] = h_for(E, Vs[nm])
    names = list(data.keys())
    return runINMF(data, k=k, lambda_=lam, nIteration=nIteration, HInit=[Hs[n].T for n in names], WInit=W,
                   VInit=[Vs[n] for n in names], seed=seed, nCores=nCores, verbose=verbose)


def optimizeSubset(scaleData, fit, cellIdx, lambda_=None, nIteration=30, seed=1, nCores=2, verbose=False):
    """Re-optimise on a subset of the cells from the current factors (R/optimizeNewParam.R:433-489).  `cellIdx` is
    {name: column indices}; re-scaling of the subset (`scaleDatasets`) is the caller's, outside the path."""
    k = _check_fit(fit, scaleData)
    lam = fit["uns"]["factorization"]["lambda"] if lambda_ is None else lambda_
    if not cellIdx or all(len(v) == 0 for v in cellIdx.values()):
        raise ValueError("No subset specified. Use either a combination of `clusterVar` and `useClusters` or "
                         "explicit `cellIdx`.")
    names = [n for n in scaleData if n in cellIdx and len(cellIdx[n]) > 0]
    data = {n: sp.csc_matrix(scaleData[n])[:, np.asarray(cellIdx[n])] for n in names}
    return runINMF(data, k=k, lambda_=lam, nIteration=nIteration,
                   HInit=[np.asarray(fit["H"][n])[:, np.asarray(cellIdx[n])].T for n in names],
                   WInit=np.asarray(fit["W"]), VInit=[np.asarray(fit["V"][n]) for n in names], seed=seed,
                   nCores=nCores, verbose=verbose)
