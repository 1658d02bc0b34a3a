"""Developer tool: random small configurations of quantileNorm / centroidAlign / online iNMF against their oracles."""
import sys, os
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import liger_b200
from liger_b200 import align
from oracle import quantile_norm_oracle as qn, centroid_align_oracle as ca, online_oracle as oo

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
bad = 0
for trial in range(int(sys.argv[2]) if len(sys.argv) > 2 else 12):
    k = int(rng.integers(2, 13))
    d = int(rng.integers(1, 4))
    sizes = [int(rng.integers(25, 400)) for _ in range(d)]
    Hs = []
    for n in sizes:
        z = rng.integers(0, k, n)
        h = rng.gamma(0.5, 0.4, (n, k)).astype(np.float32).astype(np.float64)
        h[np.arange(n), z] += rng.gamma(2.0, 1.0, n)
        h[rng.random((n, k)) < 0.2] = 0.0
        Hs.append(h)
    ref = int(rng.integers(0, d))
    nn = int(rng.integers(1, 25))
    eps = float(rng.choice([0.0, 0.5, 0.9, 2.0]))
    kw = dict(quantiles=int(rng.choice([1, 2, 7, 50])), min_cells=int(rng.choice([2, 5, 20])), n_neighbors=nn,
              max_sample=int(rng.choice([2, 10, 40, 1000])), seed=int(rng.integers(1, 100)), center=bool(rng.integers(0, 2)))
    knn = [qn.knn_ann(h, nn, eps) for h in Hs]
    out, cl = qn.quantile_norm(Hs, reference=ref, knn=knn, **kw)
    res = align.quantile_norm_hlist([h.T for h in Hs], reference=ref + 1, quantiles=kw["quantiles"], minCells=kw["min_cells"],
                                    nNeighbors=nn, maxSample=kw["max_sample"], seed=kw["seed"], center=kw["center"], eps=eps)
    e1 = float(np.abs(res["H.norm"] - np.vstack(out)).max())
    lab_ok = all((a == b).all() for a, b in zip(res["clusters"].values(), cl))
    try:
        cref = ca.centroid_align(Hs, lam=1.5)["aligned"]
        cres = align.centroid_align_hlist([h.T for h in Hs], lambda_=1.5)["aligned"]
        e2 = float(np.linalg.norm(cres - cref) / np.linalg.norm(cref))
    except Exception as ex:
        e2 = f"err {ex}"
    ok = e1 < 1e-10 and lab_ok and (isinstance(e2, float) and e2 < 1e-9)
    bad += not ok
    print(trial, "k", k, "sizes", sizes, "ref", ref, "nn", nn, "eps", eps, kw, "-> qn err", e1, "labels", lab_ok, "centroid", e2, "OK" if ok else "FAIL")
# online iNMF: ragged sizes
for trial in range(4):
    m, k = int(rng.integers(40, 120)), int(rng.integers(2, 9))
    Es = [sp.random(m, int(rng.integers(60, 300)), density=0.3, random_state=int(rng.integers(1 << 30)), format="csc") for _ in range(int(rng.integers(1, 4)))]
    mbs = int(rng.integers(20, 150))
    r = oo.online_inmf(Es, k, 3.0, max_epochs=2, minibatch_size=mbs, hals_iter=1, seed=3)
    o = liger_b200.online_inmf(Es, k=k, lambda_=3.0, maxEpochs=2, minibatchSize=mbs, seed=3)
    e = max(float(np.linalg.norm(o["W"] - r["W"]) / np.linalg.norm(r["W"])), max(float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)) for a, b in zip(o["H"], r["H"])))
    ok = e < 1e-3 and o["iterations"] == r["iterations"]
    bad += not ok
    print("online", trial, m, k, [e_.shape[1] for e_ in Es], mbs, "err", e, "OK" if ok else "FAIL")
print("failures:", bad)
sys.exit(1 if bad else 0)
