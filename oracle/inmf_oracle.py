"""TEST INFRASTRUCTURE ONLY (oracle/): FP64 numpy restatement of rliger's iNMF / UINMF ANLS path.

Nothing under liger_b200/ may import this module.  It is the checker used by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.

What it restates (all citations are into /root/reference):
  * seeded initialisation       R/integration.R:412-437  (+ R's Mersenne-Twister)
  * H / V / W block solves      R/cINMF.R:477-503        (inmf_solveH / inmf_solveV / inmf_solveW)
  * objective                   src/objErr.cpp:13-31     (objErr_i)
  * restart loop / transposes   R/integration.R:408-459
  * UINMF objective             R/integration.R:1109-1113 (block updates derived from it)
  * preprocessing that produces the path's input (only to regenerate fixtures)
        normalize               R/preprocess.R:592-608
        selectGenes             R/preprocess.R:916-990, 1005-1057, 1184-1213
        rowVars_sparse_rcpp     src/data_processing.cpp:152-172
        scaleNotCenter          R/preprocess.R:1456-1472, src/data_processing.cpp:42-58

The NNLS arithmetic itself (`RcppPlanc::bppnnls_prod`, RcppPlanc >= 2.0.0,
DESCRIPTION:60) is NOT in /root/reference; it is restated here from the
published algorithm (Kim & Park 2011, block principal pivoting).  Because the
Gram matrices are SPD the NNLS solution is unique, so any exact solver agrees.

PARITY PIN: iNMF is pinned by the reference's own golden fixture
data/pbmcPlot.rda (W, V, H of runINMF(pbmc, k=20, lambda=5, nIteration=30,
seed=1)); tests/test_oracle_golden.py checks this oracle against it.
UINMF: "parity unpinned" -- the reference holds no stored UINMF factors and
RcppPlanc draws the UINMF init internally.
"""
import numpy as np
import scipy.sparse as sp

# --------------------------------------------------------------------------
# R's default RNG (Mersenne-Twister + "Inversion"), SURVEY Appendix A.
# --------------------------------------------------------------------------


class RRandom:
    """`set.seed(seed); runif(n, lo, hi)` exactly as R does it."""

    def __init__(self, seed):
        s = np.uint32(seed & 0xFFFFFFFF)
        with np.errstate(over="ignore"):
            for _ in range(50):
                s = np.uint32(69069) * s + np.uint32(1)
            iseed = np.empty(625, dtype=np.uint32)
            for j in range(625):
                s = np.uint32(69069) * s + np.uint32(1)
                iseed[j] = s
        # FixupSeeds: dummy[0] = mti = 624  => regenerate on first draw
        self.mt = iseed[1:].copy()
        self.bitgen = np.random.MT19937()
        st = self.bitgen.state
        st["state"]["key"] = self.mt
        st["state"]["pos"] = 624
        self.bitgen.state = st

    def unif_rand(self, n):
        raw = self.bitgen.random_raw(int(n)).astype(np.float64)
        u = raw * 2.3283064365386963e-10
        # fixup(): keep strictly inside (0, 1)
        i2_32m1 = 2.328306437080797e-10
        u = np.where(u <= 0.0, 0.5 * i2_32m1, u)
        u = np.where(1.0 - u <= 0.0, 1.0 - 0.5 * i2_32m1, u)
        return u

    def runif(self, n, lo=0.0, hi=1.0):
        return lo + (hi - lo) * self.unif_rand(n)


def r_matrix_runif(rng, nrow, ncol, lo=0.0, hi=2.0):
    """matrix(abs(runif(nrow*ncol, lo, hi)), nrow, ncol) -- column-major fill."""
    return np.abs(rng.runif(nrow * ncol, lo, hi)).reshape((nrow, ncol), order="F")


def seeded_init(seed, m, k, n_list, u_list=None):
    """Init draw order of .runINMF.list (R/integration.R:412-437): W, V_1..V_d, H_1..H_d.

    For UINMF (unpinned; builder's choice stated in SURVEY 8d): W, V_1..V_d, U_1..U_d, H_1..H_d.
    Returns W (m x k), [V_i (m x k)], [U_i (u_i x k) or None], [H_i (n_i x k)].
    """
    rng = RRandom(seed)
    W = r_matrix_runif(rng, m, k)
    V = [r_matrix_runif(rng, m, k) for _ in n_list]
    U = None
    if u_list is not None:
        U = [r_matrix_runif(rng, u, k) if u > 0 else None for u in u_list]
    H = [r_matrix_runif(rng, n, k) for n in n_list]
    return W, V, U, H


# --------------------------------------------------------------------------
# NNLS by block principal pivoting (restates RcppPlanc::bppnnls_prod).
# --------------------------------------------------------------------------

def bppnnls_prod(CtC, CtB, max_iter=None, return_stats=False):
    """argmin_{X>=0} ||C X - B||_F given CtC (k x k, SPD) and CtB (k x c).

    Call sites: R/cINMF.R:481,489,501.  Columns are grouped by passive set so
    each distinct set is factorised once (pure speed-up; same arithmetic).
    Cold start F = {} (x = 0, y = -CtB) as in Kim & Park 2011, with the
    backup rule (alpha = 3) that guarantees termination.
    """
    G = np.asarray(CtC, dtype=np.float64)
    B = np.asarray(CtB, dtype=np.float64)
    if B.ndim == 1:
        B = B[:, None]
    k, c = B.shape
    if max_iter is None:
        max_iter = 5 * k + 50
    X = np.zeros((k, c))
    Y = -B.copy()
    F = np.zeros((k, c), dtype=bool)          # passive (free) set
    alpha = np.full(c, 3, dtype=np.int64)
    beta = np.full(c, k + 1, dtype=np.int64)
    pivots = np.zeros(c, dtype=np.int64)
    idx = np.arange(k)[:, None]
    for _ in range(max_iter):
        infeas = (F & (X < 0)) | (~F & (Y < 0))
        ninf = infeas.sum(axis=0)
        todo = ninf > 0
        if not todo.any():
            break
        pivots[todo] += 1
        c1 = todo & (ninf < beta)
        c2 = todo & ~c1 & (alpha >= 1)
        c3 = todo & ~c1 & ~c2
        beta[c1] = ninf[c1]
        alpha[c1] = 3
        alpha[c2] -= 1
        flip = np.zeros_like(F)
        flip[:, c1 | c2] = infeas[:, c1 | c2]
        if c3.any():
            mx = np.where(infeas[:, c3], idx, -1).max(axis=0)
            cols = np.nonzero(c3)[0]
            flip[mx, cols] = True
        F ^= flip
        cols = np.nonzero(todo)[0]
        # group by passive set
        keys = np.packbits(F[:, cols], axis=0, bitorder="little")
        _, inv = np.unique(keys.T, axis=0, return_inverse=True)
        inv = np.asarray(inv).reshape(-1)
        for g in range(inv.max() + 1):
            cc = cols[inv == g]
            f = F[:, cc[0]]
            a = ~f
            Xg = np.zeros((k, cc.size))
            Yg = np.zeros((k, cc.size))
            if f.any():
                Xg[f] = np.linalg.solve(G[np.ix_(f, f)], B[np.ix_(f, cc)])
                if a.any():
                    Yg[a] = G[np.ix_(a, f)] @ Xg[f] - B[np.ix_(a, cc)]
            else:
                Yg = -B[:, cc]
            X[:, cc] = Xg
            Y[:, cc] = Yg
    X[~F] = 0.0
    X = np.maximum(X, 0.0)
    if return_stats:
        return X, {"pivots": pivots, "passive": F, "dual": Y}
    return X


def nnls_lawson_hanson(G, b, tol=1e-12):
    """Independent active-set solver (Lawson & Hanson 1974) from normal equations.

    Used only to cross-check bppnnls_prod in tests."""
    k = b.shape[0]
    x = np.zeros(k)
    P = np.zeros(k, dtype=bool)
    w = b - G @ x
    it = 0
    while (~P).any() and w[~P].max() > tol and it < 10 * k:
        it += 1
        j = np.argmax(np.where(~P, w, -np.inf))
        P[j] = True
        while True:
            s = np.zeros(k)
            s[P] = np.linalg.solve(G[np.ix_(P, P)], b[P])
            if s[P].min() > 0:
                break
            neg = P & (s <= 0)
            alpha = np.min(x[neg] / (x[neg] - s[neg]))
            x = x + alpha * (s - x)
            P &= ~(np.abs(x) < 1e-15) | ~neg
            P[neg & (np.abs(x) < 1e-13)] = False
        x = s
        w = b - G @ x
    return x


# --------------------------------------------------------------------------
# The three block solves + objective (R/cINMF.R:477-503, src/objErr.cpp:13-31).
# H_i is n_i x k ("cell x factor") inside the loop, as in RcppPlanc.
# --------------------------------------------------------------------------

def solve_H(W, V, E, lam, U=None, P=None):
    """inmf_solveH (R/cINMF.R:477-483); UINMF adds the unshared block."""
    WV = W + V
    CtC = WV.T @ WV + lam * (V.T @ V)
    CtB = np.asarray((E.T @ WV).T)                       # k x n  == t(WV) %*% E
    if U is not None:
        CtC = CtC + (1.0 + lam) * (U.T @ U)
        CtB = CtB + np.asarray((P.T @ U).T)
    return bppnnls_prod(CtC, CtB).T                      # n x k


def solve_V(H, W, E, lam):
    """inmf_solveV (R/cINMF.R:485-491)."""
    HtH = H.T @ H
    CtC = (1.0 + lam) * HtH
    CtB = np.asarray((E @ H).T) - HtH @ W.T              # k x m
    return bppnnls_prod(CtC, CtB).T


def solve_U(H, P, lam):
    """[V;U] solve restricted to the unshared rows: CtB_U = H^T P^T."""
    HtH = H.T @ H
    CtC = (1.0 + lam) * HtH
    CtB = np.asarray((P @ H).T)
    return bppnnls_prod(CtC, CtB).T


def solve_W(Hs, Vs, Es):
    """inmf_solveW (R/cINMF.R:493-503)."""
    k = Vs[0].shape[1]
    m = Vs[0].shape[0]
    CtC = np.zeros((k, k))
    CtB = np.zeros((k, m))
    for H, V, E in zip(Hs, Vs, Es):
        HtH = H.T @ H
        CtC += HtH
        CtB += np.asarray((E @ H).T) - HtH @ V.T
    return bppnnls_prod(CtC, CtB).T


def objerr_i(H, W, V, E, lam):
    """objErr_i (src/objErr.cpp:13-31).  H is k x n here, exactly as in the C++ signature."""
    L = W + V
    sqnormE = float(E.multiply(E).sum()) if sp.issparse(E) else float((E * E).sum())
    LtL = L.T @ L
    HtH = H @ H.T
    TrLtLHtH = np.trace(LtL @ HtH)
    EtL = E.T @ L
    TrHtEtL = np.trace(H @ EtL)
    VtV = V.T @ V
    TrVtVHtH = np.trace(VtV @ HtH)
    return sqnormE - 2.0 * TrHtEtL + TrLtLHtH + lam * TrVtVHtH


def objerr_u_i(H, U, P, lam):
    """Unshared-block contribution ||P - U H||^2 + lam ||U H||^2 (R/integration.R:1109-1113)."""
    sqnormP = float(P.multiply(P).sum())
    HtH = H @ H.T
    UtU = U.T @ U
    return sqnormP - 2.0 * np.trace(H @ (P.T @ U)) + (1.0 + lam) * np.trace(UtU @ HtH)


def inmf(Es, k, lam, niter, Winit, Vinit, Hinit=None, trajectory=False):
    """Restates RcppPlanc::inmf as called at R/integration.R:438-441.

    Es: list of scipy CSC (m x n_i).  Order per iteration: H_i for all i, V_i
    for all i (old W), then W (new V_i) -- confirmed by the golden fixture.
    Returns dict(H=[n_i x k], V=[m x k], W=m x k, objErr, traj=[...]).
    """
    lam = float(lam)
    W = np.array(Winit, dtype=np.float64)
    Vs = [np.array(v, dtype=np.float64) for v in Vinit]
    Hs = [None] * len(Es)
    traj = []
    for _ in range(niter):
        for i, E in enumerate(Es):
            Hs[i] = solve_H(W, Vs[i], E, lam)
        for i, E in enumerate(Es):
            Vs[i] = solve_V(Hs[i], W, E, lam)
        W = solve_W(Hs, Vs, Es)
        if trajectory:
            traj.append(sum(objerr_i(Hs[i].T, W, Vs[i], Es[i], lam) for i in range(len(Es))))
    if niter == 0:
        Hs = [np.array(h, dtype=np.float64) for h in Hinit]
    obj = sum(objerr_i(Hs[i].T, W, Vs[i], Es[i], lam) for i in range(len(Es)))
    return {"H": Hs, "V": Vs, "W": W, "objErr": obj, "traj": traj}


def uinmf(Es, Ps, k, lams, niter, Winit, Vinit, Uinit, trajectory=False):
    """Restates RcppPlanc::uinmf (call site R/integration.R:1285-1287) from the documented
    objective (R/integration.R:1109-1113).  PARITY UNPINNED (see module header).

    Ps[i] is u_i x n_i CSC or None.  lams: scalar or per-dataset vector.
    """
    d = len(Es)
    lams = np.broadcast_to(np.asarray(lams, dtype=np.float64), (d,)).copy()
    W = np.array(Winit, dtype=np.float64)
    Vs = [np.array(v, dtype=np.float64) for v in Vinit]
    Us = [None if (Ps[i] is None) else np.array(Uinit[i], dtype=np.float64) for i in range(d)]
    Hs = [None] * d
    traj = []

    def total():
        t = 0.0
        for i in range(d):
            t += objerr_i(Hs[i].T, W, Vs[i], Es[i], lams[i])
            if Ps[i] is not None:
                t += objerr_u_i(Hs[i].T, Us[i], Ps[i], lams[i])
        return t

    for _ in range(niter):
        for i in range(d):
            Hs[i] = solve_H(W, Vs[i], Es[i], lams[i], Us[i], Ps[i])
        for i in range(d):
            Vs[i] = solve_V(Hs[i], W, Es[i], lams[i])
            if Ps[i] is not None:
                Us[i] = solve_U(Hs[i], Ps[i], lams[i])
        W = solve_W(Hs, Vs, Es)
        if trajectory:
            traj.append(total())
    return {"H": Hs, "V": Vs, "W": W, "U": Us, "objErr": total(), "traj": traj}


def run_inmf_list(Es, k=20, lam=5.0, niter=30, n_random_starts=1, seed=1, trajectory=False):
    """.runINMF.list (R/integration.R:370-461): seeded init, restarts, best-by-objErr, t(H).

    Note: `%||%` keeps restart 1's init for every later restart."""
    m = Es[0].shape[0]
    n_list = [E.shape[1] for E in Es]
    if min(n_list) < k:
        raise ValueError("Number of factors (k) should be less than the number of cells in the smallest dataset")
    if m < k:
        raise ValueError("Number of factors (k) should be less than the number of shared features")
    best = None
    W0 = V0 = H0 = None
    for i in range(n_random_starts):
        W, V, _, H = seeded_init(seed + i, m, k, n_list)
        if W0 is None:
            W0, V0, H0 = W, V, H
        out = inmf(Es, k, lam, niter, W0, V0, H0, trajectory=trajectory)
        if best is None or out["objErr"] < best["objErr"]:
            best = out
    best = dict(best)
    best["H"] = [h.T for h in best["H"]]          # k x n_i  (R/integration.R:453)
    return best


# --------------------------------------------------------------------------
# Preprocessing restatement (only what is needed to regenerate scaleData).
# --------------------------------------------------------------------------

def normalize(raw):
    """normalize.dgCMatrix (R/preprocess.R:592-608): divide each column by its sum."""
    raw = sp.csc_matrix(raw, dtype=np.float64)
    cs = np.asarray(raw.sum(axis=0)).ravel()
    out = raw.copy()
    out.data = raw.data / np.repeat(cs, np.diff(raw.indptr))
    return out


def row_vars_sparse(x, means, ncol):
    """rowVars_sparse_rcpp (src/data_processing.cpp:152-172)."""
    x = sp.csr_matrix(x)
    nrow = x.shape[0]
    rows = np.repeat(np.arange(nrow), np.diff(x.indptr))
    diff = x.data - means[rows]
    v = np.bincount(rows, weights=diff * diff, minlength=nrow)
    nz = np.bincount(rows, minlength=nrow).astype(np.float64)
    v += (x.shape[1] - nz) * means * means
    return v / (ncol - 1.0)


def select_genes_one(norm, genes, shared_mask, nUMI, thresh=0.1, alpha=0.99):
    """.selectGenes.ligerDataset + .selectGenes.withMetric (R/preprocess.R:1005-1057, 1184-1213)."""
    from scipy.stats import norm as _norm
    means = np.asarray(norm.mean(axis=1)).ravel()
    vars_ = row_vars_sparse(norm, means, norm.shape[1])
    means = means[shared_mask]
    vars_ = vars_[shared_mask]
    g = [x for x, s in zip(genes, shared_mask) if s]
    nolan = np.mean(1.0 / nUMI)
    a_corr = alpha / norm.shape[0]
    upper = means + _norm.ppf(1.0 - a_corr / 2.0) * np.sqrt(means * nolan / norm.shape[1])
    with np.errstate(divide="ignore", invalid="ignore"):
        base = np.log10(means * nolan)
        keep = (vars_ / nolan > upper) & (np.log10(vars_) > base + thresh)
    return [x for x, s in zip(g, keep) if s]


def scale_not_center(norm, genes, features):
    """scaleNotCenter.dgCMatrix (R/preprocess.R:1456-1472) + scaleNotCenter_byRow_rcpp
    (src/data_processing.cpp:42-58): rows in `features` order, each divided by sqrt(sum(x^2)/(n-1))."""
    pos = {g: i for i, g in enumerate(genes)}
    rows = np.array([pos[f] for f in features], dtype=np.int64)
    sub = sp.csr_matrix(norm)[rows, :]
    ncol = sub.shape[1]
    ss = np.asarray(sub.multiply(sub).sum(axis=1)).ravel() / (ncol - 1)
    rms = np.sqrt(ss)
    inv = np.where(rms == 0, 0.0, 1.0 / np.where(rms == 0, 1.0, rms))
    out = sp.csr_matrix(sub)
    out.data = out.data / np.repeat(np.where(rms == 0, 1.0, rms), np.diff(out.indptr))
    out.data = out.data * np.repeat((inv != 0).astype(np.float64), np.diff(out.indptr))
    out = sp.csc_matrix(out)
    out.eliminate_zeros()
    out.sort_indices()
    return out


def default_pipeline(raws, genes_list, nUMIs, combine="union", use_datasets=None):
    """normalize -> selectGenes (default args) -> scaleNotCenter for a list of raw matrices.

    Returns (features, [scaleData CSC m x n_i])."""
    norms = [normalize(r) for r in raws]
    shared = set(genes_list[0])
    for g in genes_list[1:]:
        shared &= set(g)
    sel = []
    idxs = range(len(raws)) if use_datasets is None else use_datasets
    for i in idxs:
        mask = [g in shared for g in genes_list[i]]
        sel.append(select_genes_one(norms[i], genes_list[i], mask, nUMIs[i]))
    if combine == "union":
        feats = []
        seen = set()
        for s in sel:
            for g in s:
                if g not in seen:
                    seen.add(g)
                    feats.append(g)
    else:
        feats = [g for g in sel[0] if all(g in set(s) for s in sel[1:])]
    scaled = [scale_not_center(norms[i], genes_list[i], feats) for i in range(len(raws))]
    return feats, scaled, sel
