"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the reference's quantileNorm, the step that consumes the H_i of the
hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module; the product never does.

Follows, line by line:
    .quantileNorm.HList          R/integration.R:1621-1711   (driver: clusters, kNN refinement, quantile warp)
    max_factor_rcpp              src/quantile_norm.cpp:38-67  (column scaling, arg-max factor)
    cluster_vote_rcpp            src/quantile_norm.cpp:8-35   (majority vote over the kNN, IN PLACE: the label vector the
                                                               routine walks is the one it overwrites, so cell i sees the new
                                                               labels of its neighbours j < i)
    stats::quantile (type 7), stats::approxfun(rule = 2, ties = mean), base::sample (R >= 3.6 "Rejection" sampling)
                                                              restated from their documented algorithms

Pinning (what the reference itself holds for this path): data/pbmcPlot.rda ships `H.norm` = quantileNorm(pbmcPlot) with
the defaults recorded in its command log (quantiles 50, reference "ctrl", minCells 20, nNeighbors 20, eps 0.9, refineKNN
TRUE, maxSample 1000, seed 1; RANN 2.6.1).  tests/test_oracle.py checks this restatement against it.  One ingredient is NOT
restatable: RANN::nn2(eps = 0.9) is ANN's (1 + eps)-approximate kd-tree search, whose answer depends on that library's tree
layout.  The restatement (and the CUDA path) use the EXACT k nearest neighbours, which satisfy the approximate contract;
cells whose approximate neighbour list differs can vote differently, and the test reports how many of the golden rows are
reproduced to round-off and bounds the rest.
"""
import math

import numpy as np


# ------------------------------------------------------------------------------------------------------------------
# base::sample(x, size) without replacement, R >= 3.6.0 (sample.kind = "Rejection"): R_unif_index by rejection on
# ceil(log2(n)) random bits assembled from 16-bit chunks of unif_rand(); partial Fisher-Yates on 0..n-1
# ------------------------------------------------------------------------------------------------------------------
def r_unif_index(rng, dn):
    if dn <= 0:
        return 0
    bits = int(math.ceil(math.log2(dn)))
    while True:
        v = 0
        n = 0
        while n <= bits:
            v1 = int(math.floor(rng.unif_rand(1)[0] * 65536))
            v = 65536 * v + v1
            n += 16
        if bits < 64:
            v &= (1 << bits) - 1
        if v < dn:
            return v


def r_sample_index(rng, n, size):
    """0-based indices of sample.int(n, size) (no replacement)."""
    if size < 2:
        return np.array([r_unif_index(rng, n) for _ in range(size)], dtype=np.int64)
    x = np.arange(n, dtype=np.int64)
    out = np.empty(size, dtype=np.int64)
    nn = n
    for i in range(size):
        j = r_unif_index(rng, nn)
        out[i] = x[j]
        nn -= 1
        x[j] = x[nn]
    return out


# ------------------------------------------------------------------------------------------------------------------
def max_factor(H, dims_use=None, center=False):
    """src/quantile_norm.cpp:38-67.  H: cells x k.  Returns 1-based factor ids (-1 when no used entry is > 0)."""
    H = np.array(H, dtype=np.float64, copy=True)
    n, k = H.shape
    for i in range(k):
        if center:
            H[:, i] -= H[:, i].mean()
        sd = math.sqrt(float(np.sum(H[:, i] * H[:, i])) / (n - 1))
        H[:, i] = H[:, i] / sd
    dims = np.arange(k) if dims_use is None else np.asarray(dims_use) - 1
    sub = H[:, dims]
    best = np.argmax(sub, axis=1)                 # first maximum, like the strict '>' scan
    val = sub[np.arange(n), best]
    out = dims[best] + 1
    out[~(val > 0)] = -1
    return out.astype(np.int64)


def knn_exact(H, k):
    """Exact k nearest neighbours (Euclidean, the query set is the data set, so every point finds itself first).
    Ties in distance are broken by the smaller index."""
    H = np.asarray(H, dtype=np.float64)
    n = H.shape[0]
    k = min(k, n)
    sq = np.sum(H * H, axis=1)
    out = np.empty((n, k), dtype=np.int64)
    step = 2048
    for a in range(0, n, step):
        b = min(n, a + step)
        d2 = sq[a:b, None] + sq[None, :] - 2.0 * (H[a:b] @ H.T)
        d2[np.arange(b - a), np.arange(a, b)] = -1.0          # self first
        idx = np.lexsort((np.broadcast_to(np.arange(n), d2.shape), d2), axis=1)[:, :k]
        out[a:b] = idx
    return out


def label_codes(labels):
    """as.factor(as.character(labels)): integer codes in the order of the lexicographically sorted distinct labels."""
    levels = sorted({str(int(v)) for v in labels})
    code_of = {lv: i for i, lv in enumerate(levels)}
    return np.array([code_of[str(int(v))] for v in labels], dtype=np.int64), [int(lv) for lv in levels]


def cluster_vote(nn_idx, codes, in_place=True):
    """src/quantile_norm.cpp:8-35 on factor codes: most frequent code among the neighbours, the smallest code on ties.
    in_place=True is what the routine does (it overwrites the vector it reads)."""
    src = codes if in_place else codes.copy()
    out = codes
    for i in range(nn_idx.shape[0]):
        c = src[nn_idx[i]]
        cnt = np.bincount(c)
        out[i] = int(np.argmax(cnt))              # first maximum = smallest code
    return out


def r_seq_probs(quantiles):
    by = 1.0 / quantiles
    n = int(math.floor(1.0 / by + 1e-10))
    p = 0.0 + np.arange(n + 1, dtype=np.float64) * by   # seq.default: from + (0:n) * by, capped at `to`
    return np.minimum(p, 1.0)


def quantile7(x, probs):
    """stats::quantile(x, probs, type = 7), names dropped."""
    x = np.sort(np.asarray(x, dtype=np.float64))
    n = x.shape[0]
    index = 1.0 + max(n - 1, 0) * probs
    lo = np.floor(index).astype(np.int64)
    hi = np.ceil(index).astype(np.int64)
    qs = x[lo - 1].copy()
    xh = x[hi - 1]
    sel = (index > lo) & (xh != qs)
    h = (index - lo)[sel]
    qs[sel] = (1.0 - h) * qs[sel] + h * xh[sel]
    return qs


def approx_rule2(xq, yq, v):
    """stats::approxfun(xq, yq, rule = 2, ties = mean)(v)."""
    order = np.argsort(xq, kind="stable")
    xs, ys = np.asarray(xq)[order], np.asarray(yq)[order]
    ux, inv = np.unique(xs, return_inverse=True)
    if ux.shape[0] < xs.shape[0]:
        uy = np.zeros(ux.shape[0])
        cnt = np.zeros(ux.shape[0])
        np.add.at(uy, inv, ys)
        np.add.at(cnt, inv, 1.0)
        ys = uy / cnt
        xs = ux
    v = np.asarray(v, dtype=np.float64)
    out = np.empty_like(v)
    n = xs.shape[0]
    for t, val in enumerate(v):
        if val < xs[0]:
            out[t] = ys[0]
            continue
        if val > xs[n - 1]:
            out[t] = ys[n - 1]
            continue
        i, j = 0, n - 1
        while i < j - 1:                          # R_approxfun's bisection
            ij = (i + j) // 2
            if val < xs[ij]:
                j = ij
            else:
                i = ij
        if val == xs[j]:
            out[t] = ys[j]
        elif val == xs[i]:
            out[t] = ys[i]
        else:
            out[t] = ys[i] + (ys[j] - ys[i]) * ((val - xs[i]) / (xs[j] - xs[i]))
    return out


def quantile_norm(Hs, reference=0, quantiles=50, min_cells=20, n_neighbors=20, dims_use=None, center=False,
                  max_sample=1000, refine_knn=True, seed=1, rng=None, vote_in_place=True, knn=None):
    """.quantileNorm.HList (R/integration.R:1621-1711).  Hs: list of cells x k arrays (already transposed).
    Returns (list of normalised cells x k arrays, list of per-dataset cluster labels (1-based factor ids)).
    `rng`: an R Mersenne-Twister stream positioned after set.seed(seed) (oracle.inmf_oracle.RRandom); it is only consumed
    by sample(), whose draws matter only for clusters larger than max_sample (a full-size sample is a permutation and the
    quantiles of a permutation are those of the data) -- but every call advances the stream, so all of them are made."""
    from .inmf_oracle import RRandom
    rng = rng or RRandom(seed)
    Hs = [np.array(h, dtype=np.float64, copy=True) for h in Hs]
    labels = [max_factor(h, dims_use, center) for h in Hs]
    sizes = [h.shape[0] for h in Hs]
    codes, levels = label_codes(np.concatenate(labels))
    off = np.concatenate([[0], np.cumsum(sizes)])
    if refine_knn:
        for i, h in enumerate(Hs):
            nn = knn_exact(h, n_neighbors) if knn is None else knn[i]
            seg = codes[off[i]:off[i + 1]]          # a copy in R (subsetting), then voted in place
            seg = cluster_vote(nn, seg.copy(), in_place=vote_in_place)
            codes[off[i]:off[i + 1]] = seg
    lab = np.array(levels, dtype=np.int64)[codes]
    clusters = [lab[off[i]:off[i + 1]] for i in range(len(Hs))]
    dims = Hs[reference].shape[1]
    probs = r_seq_probs(quantiles)
    # the stream only has to be followed when some sample() is a proper subset
    need_rng = any(np.bincount(c[c > 0]).max(initial=0) > max_sample for c in clusters)
    for d in range(len(Hs)):
        for c in range(1, dims + 1):
            idx2 = np.nonzero(clusters[d] == c)[0]
            idx1 = np.nonzero(clusters[reference] == c)[0]
            n2, n1 = idx2.shape[0], idx1.shape[0]
            if n1 < min_cells or n2 < min_cells:
                continue
            for k in range(dims):
                if n2 == 1:
                    Hs[d][idx2, k] = Hs[reference][idx1, k].mean()
                    continue
                v2 = Hs[d][idx2, k]
                v1 = Hs[reference][idx1, k]
                s2 = v2[r_sample_index(rng, n2, min(n2, max_sample))] if need_rng else v2
                q2 = quantile7(s2, probs)
                s1 = v1[r_sample_index(rng, n1, min(n1, max_sample))] if need_rng else v1
                q1 = quantile7(s1, probs)
                if q1.sum() == 0 or q2.sum() == 0 or np.unique(q1).shape[0] < 2 or np.unique(q2).shape[0] < 2:
                    new = np.zeros(n2)
                else:
                    new = approx_rule2(q2, q1, v2)
                Hs[d][idx2, k] = new
    return Hs, clusters


# ------------------------------------------------------------------------------------------------------------------
# RANN::nn2(data, k, eps, treetype = "kd", searchtype = "standard")  ==  ANN 1.1.2 (the library RANN 2.6.1 wraps; not under
# /root/reference): kd-tree with bucket size 1 built by the sliding-midpoint rule, standard (depth-first) k-NN search with
# the (1 + eps) pruning test.  Restated from ANN's published algorithm (Mount & Arya, ANN Programming Manual 1.1, sections
# 2.2.2 "sliding-midpoint" and 2.3 "standard search"); the traversal order and the pruning inequality decide WHICH
# approximate neighbours come back, so they are followed exactly.
# ------------------------------------------------------------------------------------------------------------------
class _AnnTree:
    ERR = 0.001

    def __init__(self, pts):
        self.pts = np.ascontiguousarray(pts, dtype=np.float64)
        n, dim = self.pts.shape
        self.dim = dim
        self.pidx = np.arange(n, dtype=np.int64)
        self.lo = self.pts.min(axis=0).copy()
        self.hi = self.pts.max(axis=0).copy()
        # nodes: ("leaf", point index or -1) / ("split", cd, cv, lv, hv, lo_child, hi_child)
        self.nodes = []
        import sys
        sys.setrecursionlimit(max(10000, sys.getrecursionlimit()))
        self.root = self._build(0, n, self.lo.copy(), self.hi.copy())

    def _plane_split(self, a, n, d, cv):
        P, idx = self.pts, self.pidx
        l, r = 0, n - 1
        while True:
            while l < n and P[idx[a + l], d] < cv:
                l += 1
            while r >= 0 and P[idx[a + r], d] >= cv:
                r -= 1
            if l > r:
                break
            idx[a + l], idx[a + r] = idx[a + r], idx[a + l]
            l += 1
            r -= 1
        br1 = l
        r = n - 1
        while True:
            while l < n and P[idx[a + l], d] <= cv:
                l += 1
            while r >= br1 and P[idx[a + r], d] > cv:
                r -= 1
            if l > r:
                break
            idx[a + l], idx[a + r] = idx[a + r], idx[a + l]
            l += 1
            r -= 1
        return br1, l

    def _build(self, a, n, blo, bhi):
        if n <= 1:
            self.nodes.append(("leaf", int(self.pidx[a]) if n == 1 else -1))
            return len(self.nodes) - 1
        P = self.pts[self.pidx[a:a + n]]
        length = bhi - blo
        max_length = length.max()
        max_spread, cd = -1.0, 0
        for d in range(self.dim):
            if length[d] >= (1.0 - self.ERR) * max_length:
                spr = P[:, d].max() - P[:, d].min()
                if spr > max_spread:
                    max_spread, cd = spr, d
        ideal = (blo[cd] + bhi[cd]) / 2.0
        mn, mx = P[:, cd].min(), P[:, cd].max()
        cv = mn if ideal < mn else (mx if ideal > mx else ideal)
        br1, br2 = self._plane_split(a, n, cd, cv)
        if ideal < mn:
            n_lo = 1
        elif ideal > mx:
            n_lo = n - 1
        elif br1 > n // 2:
            n_lo = br1
        elif br2 < n // 2:
            n_lo = br2
        else:
            n_lo = n // 2
        lv, hv = blo[cd], bhi[cd]
        bhi[cd] = cv
        lo_child = self._build(a, n_lo, blo, bhi)
        bhi[cd] = hv
        blo[cd] = cv
        hi_child = self._build(a + n_lo, n - n_lo, blo, bhi)
        blo[cd] = lv
        self.nodes.append(("split", cd, cv, lv, hv, lo_child, hi_child))
        return len(self.nodes) - 1

    def search(self, qpt, k, eps):
        max_err = (1.0 + eps) ** 2
        keys = [math.inf] * (k + 1)
        ids = [-1] * (k + 1)
        cnt = [0]
        pts, dim = self.pts, self.dim

        def max_key():
            return keys[k - 1] if cnt[0] == k else math.inf

        def insert(kv, pid):
            i = cnt[0]
            while i > 0:
                if keys[i - 1] > kv:
                    keys[i] = keys[i - 1]
                    ids[i] = ids[i - 1]
                    i -= 1
                else:
                    break
            keys[i] = kv
            ids[i] = pid
            if cnt[0] < k:
                cnt[0] += 1

        def visit(node, box_dist):
            nd = self.nodes[node]
            if nd[0] == "leaf":
                pid = nd[1]
                if pid < 0:
                    return
                md = max_key()
                dist = 0.0
                pp = pts[pid]
                for d in range(dim):
                    t = qpt[d] - pp[d]
                    dist = dist + t * t
                    if dist > md:
                        return
                insert(dist, pid)
                return
            _, cd, cv, lv, hv, lo_child, hi_child = nd
            cut_diff = qpt[cd] - cv
            if cut_diff < 0:
                visit(lo_child, box_dist)
                box_diff = lv - qpt[cd]
                if box_diff < 0:
                    box_diff = 0.0
                bd = box_dist + (cut_diff * cut_diff - box_diff * box_diff)
                if bd * max_err < max_key():
                    visit(hi_child, bd)
            else:
                visit(hi_child, box_dist)
                box_diff = qpt[cd] - hv
                if box_diff < 0:
                    box_diff = 0.0
                bd = box_dist + (cut_diff * cut_diff - box_diff * box_diff)
                if bd * max_err < max_key():
                    visit(lo_child, bd)

        # distance from the query to the bounding box of the data
        bdist = 0.0
        for d in range(dim):
            if qpt[d] < self.lo[d]:
                t = self.lo[d] - qpt[d]
                bdist += t * t
            elif qpt[d] > self.hi[d]:
                t = qpt[d] - self.hi[d]
                bdist += t * t
        visit(self.root, bdist)
        return ids[:k]


def knn_ann(H, k, eps):
    """RANN::nn2(H, k = k, eps = eps)$nn.idx, 0-based."""
    H = np.asarray(H, dtype=np.float64)
    tree = _AnnTree(H)
    k = min(k, H.shape[0])
    return np.array([tree.search(H[i], k, eps) for i in range(H.shape[0])], dtype=np.int64)
