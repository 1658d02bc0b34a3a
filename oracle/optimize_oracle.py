"""TEST INFRASTRUCTURE (oracle): CPU restatement, in dense FP64 numpy, of the reference's re-optimisation callers of
the hot path (R/optimizeNewParam.R).  Only tests/ may import this.  PARITY UNPINNED against the reference itself: the
reference's tests for these functions assert validity and shapes only (tests/testthat/test_factorization.R:104-132);
the restatement follows the R source line by line and is the checker for liger_b200/optimize.py.

    optimize_new_k       R/optimizeNewParam.R:41-163
    optimize_new_lambda  R/optimizeNewParam.R:353-385
    optimize_new_data    R/optimizeNewParam.R:245-351  (scaled data in; normalize/scaleNotCenter are outside the path)
    optimize_subset      R/optimizeNewParam.R:433-489

fit = dict(W m x k, H [k x n_i], V [m x k], lam); Es = list of dense or sparse m x n_i.
"""
import numpy as np
import scipy.sparse as sp

from . import inmf_oracle as orc


def _d(x):
    return np.asarray(x.todense()) if sp.issparse(x) else np.asarray(x, dtype=np.float64)


def _run(Es, k, lam, niter, H, W, V, seed):
    """runINMF with explicit inits (R/integration.R:412-441): missing inits are drawn from set.seed(seed) in the order
    W, V_1..V_d, H_1..H_d, skipping the ones that were passed."""
    rng = orc.RRandom(seed)
    m = W.shape[0] if W is not None else Es[0].shape[0]
    if W is None:
        W = orc.r_matrix_runif(rng, m, k)
    if V is None:
        V = [orc.r_matrix_runif(rng, m, k) for _ in Es]
    if H is None:
        H = [orc.r_matrix_runif(rng, E.shape[1], k) for E in Es]
    out = orc.inmf([sp.csc_matrix(E) for E in Es], k, lam, niter, W, V, H)
    return {"W": out["W"], "V": out["V"], "H": [h.T for h in out["H"]], "objErr": out["objErr"], "lam": lam}


def optimize_new_k(Es, fit, k_new, niter, seed=1, lam=None):
    lam = fit["lam"] if lam is None else lam
    W, V, H = fit["W"], fit["V"], fit["H"]
    k = W.shape[1]
    g = W.shape[0]
    d = len(Es)
    if k_new == k:
        return fit
    E = [_d(e) for e in Es]
    if k_new > k:
        kd = k_new - k
        rng = orc.RRandom(seed)
        W_new = rng.runif(g * kd, 0.0, 2.0).reshape((kd, g), order="F").T          # t(matrix(runif(..), kd, g))
        V_new = [rng.runif(g * kd, 0.0, 2.0).reshape((kd, g), order="F").T for _ in range(d)]
        H_new = []
        for i in range(d):                                                          # :103-119
            L_new = W_new + V_new[i]
            L = W + V[i]
            CtC = L_new.T @ L_new + lam * V_new[i].T @ V_new[i]
            CtB = -1.0 * (L_new.T @ L @ H[i] - L_new.T @ E[i])
            H_new.append(orc.bppnnls_prod(CtC, CtB))
        V_new2 = []
        for i in range(d):                                                          # :120-133
            CtC = H_new[i] @ H_new[i].T * (1.0 + lam)
            CtB = H_new[i] @ E[i].T - H_new[i] @ H[i].T @ (W + V[i]).T - H_new[i] @ H_new[i].T @ W_new.T
            V_new2.append(orc.bppnnls_prod(CtC, CtB).T)
        V_new = V_new2
        CtC = np.zeros((kd, kd))
        CtB = np.zeros((kd, g))
        for i in range(d):                                                          # :139-149
            CtC += H_new[i] @ H_new[i].T
            CtB += H_new[i] @ E[i].T - H_new[i] @ H[i].T @ (W + V[i]).T - H_new[i] @ H_new[i].T @ V_new[i].T
        W_new = orc.bppnnls_prod(CtC, CtB).T
        Hn = [np.vstack([H[i], H_new[i]]).T for i in range(d)]
        Vn = [np.hstack([V[i], V_new[i]]) for i in range(d)]
        Wn = np.hstack([W, W_new])
    else:
        deltas = np.zeros(k)
        for i in range(d):                                                          # :141-146
            for ki in range(k):
                deltas[ki] += np.linalg.norm(np.outer(W[:, ki] + V[i][:, ki], H[i][ki, :]), "fro")
        k_use = np.argsort(-deltas, kind="stable")[:k_new]
        Wn = W[:, k_use]
        Vn = [v[:, k_use] for v in V]
        Hn = [h[k_use, :].T for h in H]
    return _run(Es, k_new, lam, niter, Hn, Wn, Vn, seed)


def optimize_new_lambda(Es, fit, lam_new, niter, seed=1):
    return _run(Es, fit["W"].shape[1], lam_new, niter, [h.T for h in fit["H"]], fit["W"], None, seed)


def _h_for(W, V, E, lam):
    L = W + V
    return orc.bppnnls_prod(L.T @ L + lam * V.T @ V, L.T @ _d(E))


def optimize_new_data(Es, fit, Es_new, use, merge, new_cells=None, niter=30, seed=1, lam=None):
    """use[i] = index of the existing dataset Es_new[i] is merged into / inherits V from."""
    lam = fit["lam"] if lam is None else lam
    W = fit["W"]
    Es2, H2, V2 = list(Es), list(fit["H"]), list(fit["V"])
    if merge:
        for E, t, idx in zip(Es_new, use, new_cells):
            E = sp.csc_matrix(E)
            H2[t] = np.hstack([H2[t], _h_for(W, V2[t], E[:, idx], lam)])
            Es2[t] = E
    else:
        for E, s in zip(Es_new, use):
            Es2.append(sp.csc_matrix(E))
            V2.append(V2[s].copy())
            H2.append(_h_for(W, V2[-1], E, lam))
    return _run(Es2, W.shape[1], lam, niter, [h.T for h in H2], W, V2, seed)


def optimize_subset(Es, fit, cell_idx, niter, seed=1, lam=None):
    lam = fit["lam"] if lam is None else lam
    Es2 = [sp.csc_matrix(E)[:, idx] for E, idx in zip(Es, cell_idx)]
    return _run(Es2, fit["W"].shape[1], lam, niter, [h[:, idx].T for h, idx in zip(fit["H"], cell_idx)], fit["W"],
                fit["V"], seed)
