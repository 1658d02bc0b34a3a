"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, FP64) of online iNMF, scenarios 1-3.  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product never does.

PARITY UNPINNED.  The reference's runOnlineINMF (R/integration.R:722-933) hands everything -- initialisation, minibatch
sampling, the HALS updates -- to RcppPlanc::onlineINMF (R/integration.R:902-910), whose source is not under /root/reference,
and the reference's tests only check shapes (tests/testthat/test_factorization.R:101-178).  What is restated here is the
PUBLISHED algorithm (Gao et al., Nat Biotechnol 2021, Algorithm 1: minibatch H solves, running sufficient statistics
A_i = H H^T / b, B_i = X H^T / b, one HALS sweep over the columns of W and V_i per minibatch, a final H solve over all
cells), with the bookkeeping of the R implementation that paper shipped (rliger 1.0 `online_iNMF`):

    set.seed(seed);  W = runif(m k, 0, 2);  V_i = the scaled data of k sampled cells (sample.int(n_i, k));  columns of W and
    V_i scaled to unit 2-norm;  A_i = 0, B_i = 0;  one permutation sample.int(n_i) per dataset and epoch
    minibatch sizes b_i = round(n_i / N * minibatchSize);  iterations t = 1 .. floor(N * maxEpochs / minibatchSize)
    A_i <- s_t A_i + H_b H_b^T / b_i,  B_i <- s_t B_i + X_b H_b^T / b_i,   s_1 = 0, s_2 = 1, s_t = (t - 2) / (t - 1);  zero
    diagonal entries of A_i are lifted to 1e-15
    W[, j]   <- nonneg(W[, j]   + sum_i (B_i[, j] - (W + V_i) A_i[, j]) / sum_i A_i[j, j])                      j = 1 .. k
    V_i[, j] <- nonneg(V_i[, j] + (B_i[, j] - (W + (1 + lambda) V_i) A_i[, j]) / ((1 + lambda) A_i[j, j]))      nonneg(x) = max(x, 1e-16)

Scenario 2 (`newDatasets`, R/integration.R:755-812): W, V_i, A_i, B_i of the already factorised datasets are given; no W draw;
the new datasets draw their V_i (sample.int(n_i, k), unit columns) and start from A_i = B_i = 0; minibatch sizes, the number
of iterations and the permutations count the NEW cells only; every iteration solves H_b and updates A_i, B_i for the new
datasets, sweeps W over the statistics of ALL datasets and V_i over the new ones; at the end H_i of the new datasets.
Scenario 3 (`projection = TRUE`): H_i = argmin_{H >= 0} ||X_i - W H||_F^2 for the new datasets, nothing else changes.
"""
import numpy as np

from . import inmf_oracle as orc
from .quantile_norm_oracle import r_sample_index


def schedule(sizes, minibatch_size, max_epochs, rng):
    """Cell indices of every minibatch: list over iterations of a list over datasets (0-based indices).  A dataset walks
    through a permutation of its cells; when a minibatch runs past the end, the remaining cells come from the head of a
    fresh permutation (drawn at that moment)."""
    N = int(sum(sizes))
    mb = [max(1, int(round(n / N * minibatch_size))) for n in sizes]
    iters = int(N * max_epochs // minibatch_size)
    perm = [r_sample_index(rng, n, n) for n in sizes]
    pos = [0] * len(sizes)
    out = []
    for _ in range(iters):
        cur = []
        for i, n in enumerate(sizes):
            take = mb[i]
            idx = []
            while take > 0:
                room = n - pos[i]
                if room == 0:
                    perm[i] = r_sample_index(rng, n, n)
                    pos[i] = 0
                    room = n
                c = min(take, room)
                idx.append(perm[i][pos[i]:pos[i] + c])
                pos[i] += c
                take -= c
            cur.append(np.concatenate(idx))
        out.append(cur)
    return out, mb


def init(Es, k, seed):
    """W, V_i of the R implementation (see the header) and the RNG stream positioned after them."""
    rng = orc.RRandom(seed)
    m = Es[0].shape[0]
    W = rng.runif(m * k, 0.0, 2.0).reshape((m, k), order="F")
    Vs = []
    for E in Es:
        cols = r_sample_index(rng, E.shape[1], k)
        Vs.append(np.asarray(E[:, cols].todense(), dtype=np.float64))
    W = W / np.sqrt((W * W).sum(axis=0))
    Vs = [V / np.where((V * V).sum(axis=0) > 0, np.sqrt((V * V).sum(axis=0)), 1.0) for V in Vs]
    return W, Vs, rng


def nonneg(x):
    return np.maximum(x, 1e-16)


def online_inmf(Es, k, lam, max_epochs=5, minibatch_size=5000, hals_iter=1, seed=1, Winit=None, Vinit=None):
    Es = [E.tocsc() for E in Es]
    sizes = [E.shape[1] for E in Es]
    m = Es[0].shape[0]
    W, Vs, rng = init(Es, k, seed)
    if Winit is not None:
        W, Vs = np.array(Winit, dtype=np.float64), [np.array(v, dtype=np.float64) for v in Vinit]
    sched, mb = schedule(sizes, minibatch_size, max_epochs, rng)
    A = [np.zeros((k, k)) for _ in Es]
    B = [np.zeros((m, k)) for _ in Es]
    for t, cur in enumerate(sched, start=1):
        s = 0.0 if t == 1 else (1.0 if t == 2 else (t - 2.0) / (t - 1.0))
        for i, E in enumerate(Es):
            Xb = E[:, cur[i]]
            Hb = orc.solve_H(W, Vs[i], Xb, lam)                       # b_i x k
            A[i] = s * A[i] + Hb.T @ Hb / mb[i]
            d = np.diag_indices(k)
            A[i][d] = np.where(A[i][d] == 0, 1e-15, A[i][d])
            B[i] = s * B[i] + np.asarray(Xb @ Hb) / mb[i]
        for _ in range(hals_iter):
            for j in range(k):
                num = np.zeros(m)
                den = 0.0
                for i in range(len(Es)):
                    num += B[i][:, j] - (W + Vs[i]) @ A[i][:, j]
                    den += A[i][j, j]
                W[:, j] = nonneg(W[:, j] + num / den)
            for j in range(k):
                for i in range(len(Es)):
                    Vs[i][:, j] = nonneg(Vs[i][:, j] + (B[i][:, j] - (W + (1.0 + lam) * Vs[i]) @ A[i][:, j]) / ((1.0 + lam) * A[i][j, j]))
    Hs = [orc.solve_H(W, Vs[i], E, lam) for i, E in enumerate(Es)]
    obj = sum(orc.objerr_i(Hs[i].T, W, Vs[i], Es[i], lam) for i in range(len(Es)))
    return {"W": W, "V": Vs, "H": Hs, "A": A, "B": B, "objErr": obj, "iterations": len(sched), "minibatch": mb}


def online_inmf_new_datasets(W, V_prev, A_prev, B_prev, Es_new, k, lam, max_epochs=5, minibatch_size=5000, hals_iter=1, seed=1):
    """Scenario 2: see the module header.  Returns the state of all datasets (previous ones first) and H of the new ones."""
    Es_new = [E.tocsc() for E in Es_new]
    sizes = [E.shape[1] for E in Es_new]
    m = Es_new[0].shape[0]
    npv = len(V_prev)
    rng = orc.RRandom(seed)
    Vn = []
    for E in Es_new:
        cols = r_sample_index(rng, E.shape[1], k)
        V = np.asarray(E[:, cols].todense(), dtype=np.float64)
        Vn.append(V / np.where((V * V).sum(axis=0) > 0, np.sqrt((V * V).sum(axis=0)), 1.0))
    sched, mb = schedule(sizes, minibatch_size, max_epochs, rng)
    W = np.array(W, dtype=np.float64)
    Vs = [np.array(v, dtype=np.float64) for v in V_prev] + Vn
    A = [np.array(a, dtype=np.float64) for a in A_prev] + [np.zeros((k, k)) for _ in Es_new]
    B = [np.array(b, dtype=np.float64) for b in B_prev] + [np.zeros((m, k)) for _ in Es_new]
    d = len(Vs)
    for t, cur in enumerate(sched, start=1):
        s = 0.0 if t == 1 else (1.0 if t == 2 else (t - 2.0) / (t - 1.0))
        for q, E in enumerate(Es_new):
            i = npv + q
            Xb = E[:, cur[q]]
            Hb = orc.solve_H(W, Vs[i], Xb, lam)
            A[i] = s * A[i] + Hb.T @ Hb / mb[q]
            dg = np.diag_indices(k)
            A[i][dg] = np.where(A[i][dg] == 0, 1e-15, A[i][dg])
            B[i] = s * B[i] + np.asarray(Xb @ Hb) / mb[q]
        for _ in range(hals_iter):
            for j in range(k):
                num = np.zeros(m)
                den = 0.0
                for i in range(d):
                    num += B[i][:, j] - (W + Vs[i]) @ A[i][:, j]
                    den += A[i][j, j]
                W[:, j] = nonneg(W[:, j] + num / den)
            for j in range(k):
                for i in range(npv, d):
                    Vs[i][:, j] = nonneg(Vs[i][:, j] + (B[i][:, j] - (W + (1.0 + lam) * Vs[i]) @ A[i][:, j]) / ((1.0 + lam) * A[i][j, j]))
    Hs = [orc.solve_H(W, Vs[npv + q], E, lam) for q, E in enumerate(Es_new)]
    obj = sum(orc.objerr_i(Hs[q].T, W, Vs[npv + q], Es_new[q], lam) for q in range(len(Es_new)))
    return {"W": W, "V": Vs, "H": Hs, "A": A, "B": B, "objErr": obj, "iterations": len(sched), "minibatch": mb}


def project(W, Es_new, lam=5.0):
    """Scenario 3: H of the new datasets against W alone (V = 0, so the penalty term vanishes)."""
    W = np.asarray(W, dtype=np.float64)
    return [orc.solve_H(W, np.zeros_like(W), E.tocsc(), lam) for E in Es_new]
