"""Mirror of the three R-level block solves of the reference (R/cINMF.R:477-503) and of the "consensus fixed"
refinement loop of runCINMF that is built from them (R/cINMF.R:392-407).  Each solve is host algebra for the normal
equations followed by RcppPlanc::bppnnls_prod -- here lgr_bppnnls_prod, i.e. the NNLS runs on the GPU -- and the
objective is objErr_i (src/objErr.cpp:13-31 -> lgr_objerr_i).

    inmf_solveH(H, W, V, E, lambda)     -> n x k      R/cINMF.R:477-483
    inmf_solveV(H, W, V, E, lambda)     -> m x k      R/cINMF.R:485-491
    inmf_solveW(Hs, W, Vs, Es, lambda)  -> m x k      R/cINMF.R:493-503
    refine_fixed_W(Es, W, Vs, lambda, nIteration, seed) : the loop of R/cINMF.R:392-407 (W fixed to the consensus,
        H_i / V_i alternate 2 * nIteration times), returning dict(H, V, W, objErr)
"""
import numpy as np
import scipy.sparse as sp

from . import api


def _dense(x):
    return np.asarray(x.todense()) if sp.issparse(x) else np.asarray(x)


def inmf_solveH(H, W, V, E, lambda_, nCores=2):
    WV = W + V
    CtC = WV.T @ WV + lambda_ * (V.T @ V)
    CtB = _dense((E.T @ WV).T)
    return api.bppnnls_prod(CtC, CtB, nCores=nCores).T          # cell x factor


def inmf_solveV(H, W, V, E, lambda_, nCores=2):
    HtH = H.T @ H
    CtC = (1.0 + lambda_) * HtH
    CtB = _dense((E @ H).T) - HtH @ W.T
    return api.bppnnls_prod(CtC, CtB, nCores=nCores).T


def inmf_solveW(Hs, W, Vs, Es, lambda_, nCores=2):
    k, m = Vs[0].shape[1], Vs[0].shape[0]
    CtC = np.zeros((k, k))
    CtB = np.zeros((k, m))
    for H, V, E in zip(Hs, Vs, Es):
        HtH = H.T @ H
        CtC += HtH
        CtB += _dense((E @ H).T) - HtH @ V.T
    return api.bppnnls_prod(CtC, CtB, nCores=nCores).T


def refine_fixed_W(Es, W, Vs, lambda_=5.0, nIteration=30, seed=1, nCores=2):
    """ANLS with the consensus W fixed (R/cINMF.R:386-407).  The H inits are drawn (and immediately overwritten by the
    first H solve) exactly like the reference does, so the RNG stream stays aligned with R."""
    k = W.shape[1]
    rng = api.RRandom(seed)
    Hs = [rng.runif(E.shape[1] * k, 0.0, 2.0).reshape((E.shape[1], k), order="F") for E in Es]
    Vs = [np.asarray(v, dtype=np.float64) for v in Vs]
    for _ in range(2 * nIteration):
        for i, E in enumerate(Es):
            Hs[i] = inmf_solveH(None, W, Vs[i], E, lambda_, nCores)
        for i, E in enumerate(Es):
            Vs[i] = inmf_solveV(Hs[i], W, None, E, lambda_, nCores)
    obj = sum(api.objErr_i(Hs[i].T, W, Vs[i], Es[i], lambda_) for i in range(len(Es)))
    return {"H": Hs, "V": Vs, "W": W, "objErr": obj}
