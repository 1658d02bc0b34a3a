"""Developer check (GPU box): pbmc golden parity through the C ABI."""
import sys, os, time
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import liger_b200
from oracle import inmf_oracle as orc

g = np.load("tests/golden/pbmc_golden.npz")
Es = [sp.csc_matrix((g[f"scale_{n}_data"], g[f"scale_{n}_indices"], g[f"scale_{n}_indptr"]), shape=tuple(g[f"scale_{n}_shape"])) for n in ("ctrl", "stim")]
rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)

# 1. bppnnls_prod
rng = np.random.default_rng(0)
C = rng.random((50, 20)); G = C.T @ C; B = C.T @ rng.random((50, 700))
X = liger_b200.bppnnls_prod(G, B)
Xo = orc.bppnnls_prod(G, B)
print("bppnnls_prod rel err", rel(X, Xo), "pattern equal", np.array_equal(X > 0, Xo > 0))

# 2. inmf on pbmc
t = time.time()
out = liger_b200.run_inmf_list(Es, k=20, lambda_=5.0, nIteration=30, seed=1, trajectory=True)
print("run_inmf_list time", time.time() - t, out["timings"])
rows = g["golden_rows"]
print("W", rel(out["W"][rows], g["golden_W"]))
for i, n in enumerate(("ctrl", "stim")):
    print(n, "V", rel(out["V"][i][rows], g[f"golden_V_{n}"]), "H", rel(out["H"][i], g[f"golden_H_{n}"]),
          "zero-pattern mismatches", int(((out["H"][i] == 0) != (g[f"golden_H_{n}"] == 0)).sum()))
print("traj", out["traj"][:5], out["traj"][-1], "objErr", out["objErr"], "oracle 36105.069129349")
