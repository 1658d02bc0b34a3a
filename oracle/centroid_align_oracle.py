"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, FP64) of the reference's centroidAlign.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product never does.

Follows:
    .centroidAlign.Hraw            R/integration.R:2015-2084
    safe_scale                     src/data_processing.cpp:61-85
    normalize_byCol_dense_rcpp     src/data_processing.cpp:10-20
    moe_correct_ridge_cpp          src/centoidAlign.cpp:16-85   (restated with dense algebra, cluster by cluster)

PARITY UNPINNED: the reference holds no stored output of centroidAlign (tests/testthat/test_factorization.R:216-228 checks
the error message, the shape of H.norm and the class of the diagnostics only), so this restatement is anchored on the
source lines above alone.
"""
import numpy as np


def safe_scale(x, center, scale):
    x = np.array(x, dtype=np.float64, copy=True)
    means = x.mean(axis=0)
    if center:
        x -= means
    if scale:
        for i in range(x.shape[1]):
            sd = np.sqrt(np.sum(x[:, i] * x[:, i]) / (x.shape[0] - 1))
            x[:, i] = 0.0 if sd == 0 else x[:, i] / sd
    return x


def normalize_by_col(x):
    x = np.array(x, dtype=np.float64, copy=True)
    sums = x.sum(axis=0)
    for i in range(x.shape[1]):
        x[:, i] = 0.0 if sums[i] == 0 else x[:, i] / sums[i]
    return x


def moe_correct_ridge(Z_orig, R, lam, batch, B):
    """Z_orig: D x N, R: K x N, lam: B, batch: N integer codes (Phi = one-hot of them)."""
    D, N = Z_orig.shape
    K = R.shape[0]
    Phi_moe = np.zeros((B + 1, N))
    Phi_moe[0] = 1.0
    Phi_moe[batch + 1, np.arange(N)] = 1.0
    lambda_mat = np.diag(np.concatenate([[0.0], np.asarray(lam, dtype=np.float64)]))
    Z_corr = Z_orig.copy()
    for k in range(K):
        Phi_Rk = Phi_moe * R[k][None, :]
        inv_cov = np.linalg.inv(Phi_Rk @ Phi_moe.T + lambda_mat)
        Z_tmp = Z_orig * R[k][None, :]
        W = np.outer(inv_cov[:, 0], Z_tmp.sum(axis=1))
        for b in range(B):
            W += np.outer(inv_cov[:, b + 1], Z_tmp[:, batch == b].sum(axis=1))
        W[0] = 0.0
        Z_corr -= W.T @ Phi_Rk
    return Z_corr


def centroid_align(Hs, lam=1.0, dims_use=None, scale_emb=True, center_emb=True, scale_cluster=False, center_cluster=False,
                   shift=False, diagnosis=False):
    """Hs: list of cells x k arrays (one per dataset, in order).  Returns dict(aligned = N x kd [, raw/Z/R which.max])."""
    obj = np.vstack([np.asarray(h, dtype=np.float64) for h in Hs])
    batch = np.concatenate([np.full(h.shape[0], i, dtype=np.int64) for i, h in enumerate(Hs)])
    B = len(Hs)
    lam = np.full(B, float(lam)) if np.ndim(lam) == 0 else np.asarray(lam, dtype=np.float64)
    if dims_use is not None:
        obj = obj[:, np.asarray(dims_use) - 1]
    out = {}
    if diagnosis:
        out["raw_which.max"] = np.argmax(obj, axis=1) + 1
    Z = safe_scale(obj, center_emb, scale_emb).T
    R = safe_scale(obj, center_cluster, scale_cluster).T
    if shift:
        R = R - R.min()
    if (R < 0).any():
        raise ValueError("Negative values found prior to normalizing the cluster assignment probablity distribution")
    R = normalize_by_col(R)
    Zc = moe_correct_ridge(Z, R, lam, batch, B).T
    out["aligned"] = Zc
    if diagnosis:
        out["Z_which.max"] = np.argmax(Zc, axis=1) + 1
        out["R_which.max"] = np.argmax(R, axis=0) + 1
    return out
