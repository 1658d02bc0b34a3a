"""Mirror of the R wrappers around the native iNMF / UINMF calls.

    run_inmf_list   <- .runINMF.list   (reference R/integration.R:370-461)
    run_uinmf_list  <- .runUINMF.list  (reference R/integration.R:1254-1310)
    runINMF / runUINMF: thin entry points over {name: scaleData} dicts mirroring
                       runINMF.liger (R/integration.R:230-284) / runUINMF.liger (:1206-1250)

Same argument names and meaning, same validation errors, same seeded initialisation
(R's Mersenne-Twister stream; draw order W, V_1..V_d, H_1..H_d), same restart selection
(best objErr; note `%||%` keeps restart 1's inits for later restarts, R/integration.R:413-428),
same output orientation (H transposed to k x n_i, R/integration.R:453).
"""
import numpy as np

from . import api


def _r_matrix(rng, nrow, ncol):
    # matrix(abs(runif(nrow*ncol, 0, 2)), nrow, ncol): column-major fill
    return np.abs(rng.runif(nrow * ncol, 0.0, 2.0)).reshape((nrow, ncol), order="F")


def run_inmf_list(objectList, k=20, lambda_=5.0, nIteration=30, nRandomStarts=1, HInit=None, WInit=None, VInit=None,
                  seed=1, nCores=2, verbose=False, trajectory=False):
    mats = list(objectList.values()) if isinstance(objectList, dict) else list(objectList)
    names = list(objectList.keys()) if isinstance(objectList, dict) else None
    m = mats[0].shape[0]
    for x in mats:
        if x.shape[0] != m:
            raise ValueError("All datasets must share the same features (rows)")
    n_list = [x.shape[1] for x in mats]
    if min(n_list) < k:
        raise ValueError(f"Number of factors (k={k}) should be less than the number of cells in the smallest "
                         f"dataset ({min(n_list)}).")
    if m < k:
        raise ValueError(f"Number of factors (k={k}) should be less than the number of shared features ({m}).")
    best = None
    best_seed = seed
    for i in range(nRandomStarts):
        rng = api.RRandom(seed + i)
        if WInit is None:
            WInit = _r_matrix(rng, m, k)
        if VInit is None:
            VInit = [_r_matrix(rng, m, k) for _ in mats]
        if HInit is None:
            HInit = [_r_matrix(rng, n, k) for n in n_list]
        out = api.inmf(mats, k=k, lambda_=lambda_, niter=nIteration, Hinit=HInit, Vinit=VInit, Winit=WInit,
                       nCores=nCores, verbose=verbose, trajectory=trajectory)
        # strict `<` in the reference (R/integration.R:442-446).  The device sums in a run-dependent order, so equal problems
        # tie to ~1e-9 instead of exactly: a later restart has to win by more than that noise, which keeps the reference's
        # outcome (the first of equal restarts) deterministic
        if best is None or out["objErr"] < best["objErr"] * (1.0 - 1e-7):
            best = out
            best_seed = seed + i
    res = dict(best)
    res["H"] = [h.T.copy(order="F") for h in best["H"]]   # k x n_i
    res["bestSeed"] = best_seed
    if names is not None:
        res["H"] = dict(zip(names, res["H"]))
        res["V"] = dict(zip(names, res["V"]))
    return res


def run_uinmf_list(objectList, unsharedList, k=20, lambda_=5.0, nIteration=30, nRandomStarts=1, seed=1, nCores=2,
                   verbose=False, trajectory=False):
    mats = list(objectList.values()) if isinstance(objectList, dict) else list(objectList)
    names = list(objectList.keys()) if isinstance(objectList, dict) else None
    if isinstance(unsharedList, dict):
        if names is None:
            raise ValueError("unsharedList is a dict but objectList is not: name the datasets in both, or pass two lists")
        uns = [unsharedList.get(nm) for nm in names]
    else:
        uns = list(unsharedList)
    if len(uns) != len(mats):
        raise ValueError("unsharedList must have one entry (or None) per dataset")
    m = mats[0].shape[0]
    n_list = [x.shape[1] for x in mats]
    if min(n_list) < k:
        raise ValueError(f"Number of factors (k={k}) should be less than the number of cells in the smallest "
                         f"dataset ({min(n_list)}).")
    if all(u is None for u in uns):
        raise ValueError("No scaled data for unshared feature found. Run `selectGenes()` with "
                         "`useUnsharedDatasets` specified, and then `scaleNotCenter()`.")
    u_list = [0 if u is None else u.shape[0] for u in uns]
    best = None
    best_seed = None
    for i in range(nRandomStarts):
        seed = seed + i            # cumulative, as in R/integration.R:1283
        rng = api.RRandom(seed)
        W0 = _r_matrix(rng, m, k)
        V0 = [_r_matrix(rng, m, k) for _ in mats]
        U0 = [(_r_matrix(rng, u, k) if u > 0 else None) for u in u_list]
        _ = [rng.runif(n * k, 0.0, 2.0) for n in n_list]      # H draws (unused: H is solved first)
        out = api.uinmf(mats, uns, k=k, lambda_=lambda_, niter=nIteration, nCores=nCores, verbose=verbose,
                        Winit=W0, Vinit=V0, Uinit=U0, trajectory=trajectory)
        if best is None or out["objErr"] < best["objErr"]:
            best = out
            best_seed = seed
    res = dict(best)
    res["H"] = [h.T.copy(order="F") for h in best["H"]]
    res["bestSeed"] = best_seed
    if names is not None:
        res["H"] = dict(zip(names, res["H"]))
        res["V"] = dict(zip(names, res["V"]))
        res["U"] = {nm: u for nm, u in zip(names, res["U"]) if u is not None}
    return res


def runINMF(scaleData, k=20, lambda_=5.0, nIteration=30, nRandomStarts=1, HInit=None, WInit=None, VInit=None, seed=1,
            nCores=2, verbose=False):
    """runINMF.liger on a {dataset: scaleData} dict: returns dict(W, H={...k x n_i}, V={...}, objErr,
    uns=dict(factorization=dict(k, lambda)))  (R/integration.R:230-284)."""
    if any(v is None for v in scaleData.values()):
        raise ValueError("Scaled data not available. Run `scaleNotCenter()` first.")
    out = run_inmf_list(scaleData, k, lambda_, nIteration, nRandomStarts, HInit, WInit, VInit, seed, nCores, verbose)
    out["uns"] = {"factorization": {"k": k, "lambda": lambda_}}
    return out


def runUINMF(scaleData, scaleUnsharedData, k=20, lambda_=5.0, nIteration=30, nRandomStarts=1, seed=1, nCores=2,
             verbose=False):
    """runUINMF.liger on dicts of scaleData / scaleUnsharedData (R/integration.R:1206-1250)."""
    if any(v is None for v in scaleData.values()):
        raise ValueError("Scaled data not available. Run `scaleNotCenter()` first.")
    out = run_uinmf_list(scaleData, scaleUnsharedData, k, lambda_, nIteration, nRandomStarts, seed, nCores, verbose)
    out["uns"] = {"factorization": {"k": k, "lambda": lambda_}}
    return out


def runOnlineINMF(scaleData, k=20, lambda_=5.0, newDatasets=None, projection=False, maxEpochs=5, HALSiter=1,
                  minibatchSize=5000, HInit=None, WInit=None, VInit=None, AInit=None, BInit=None, seed=1, nCores=2,
                  verbose=False):
    """runOnlineINMF.liger (R/integration.R:722-933) over {dataset: scaleData}.  Scenario 1: online iNMF from scratch.
    Scenario 2 (`newDatasets` = {name: scaleData of the new datasets}, already normalised and scaled -- the R front end does
    that at :806-807): `scaleData` names the already factorised datasets, whose W / V / A / B (and H) come in as WInit, VInit,
    AInit, BInit (HInit) dicts; everything is returned for old and new datasets.  Scenario 3 (`projection` = True): only H
    of the new datasets, against WInit.  Returns dict(W, H={name: k x n_i}, V, A, B, objErr, uns)."""
    names = list(scaleData.keys())
    pick = lambda init, ns: None if init is None else [init[n] if isinstance(init, dict) else v for n, v in zip(ns, init)]
    if newDatasets is None:
        if projection:
            raise ValueError("projection needs newDatasets")
        if any(v is None for v in scaleData.values()):
            raise ValueError("Scaled data not available. Run `scaleNotCenter()` first.")
        res = api.online_inmf([scaleData[n] for n in names], k=k, lambda_=lambda_, maxEpochs=maxEpochs,
                              minibatchSize=minibatchSize, HALSiter=HALSiter, Winit=WInit, Vinit=pick(VInit, names), seed=seed)
        out_names = names
    else:
        new_names = list(newDatasets.keys())
        if any(n in scaleData for n in new_names):
            raise ValueError("Names of `newDatasets` overlap with existing datasets.")
        if WInit is None or (not projection and any(x is None for x in (VInit, AInit, BInit))):
            raise ValueError("Cannot find complete online iNMF result for current datasets. "
                             "Please run `runOnlineINMF()` without `newDataset` first")
        res = api.online_inmf([scaleData[n] for n in names], k=k, lambda_=lambda_, maxEpochs=maxEpochs,
                              minibatchSize=minibatchSize, HALSiter=HALSiter, Winit=WInit, Vinit=pick(VInit, names),
                              Ainit=pick(AInit, names), Binit=pick(BInit, names),
                              Hinit=None if HInit is None else [np.asarray(h).T for h in pick(HInit, names)], seed=seed,
                              newDatasets=[newDatasets[n] for n in new_names], projection=projection)
        if projection:   # scenario 3: only H of the new datasets (R/integration.R:839-849)
            return {"H": {n: h.T for n, h in zip(new_names, res["H"])}, "uns": {"factorization": {"k": k, "lambda": lambda_}}}
        out_names = names + new_names
    names = out_names
    out = {"W": res["W"], "objErr": res["objErr"], "uns": {"factorization": {"k": k, "lambda": lambda_}}}
    out["H"] = {n: (None if h is None else h.T) for n, h in zip(names, res["H"])}   # k x n_i, as the R wrapper stores them (:924)
    for key in ("V", "A", "B"):
        out[key] = dict(zip(names, res[key]))
    return out
