"""quantileNorm of the resident H at full size: C3 factorised for a few iterations, then lgr_ctx_quantile_norm (developer tool).

    python scripts/align_at_scale.py [config] [n_per] [niter]
"""
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, liger_b200
from liger_b200 import align
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
cfg = bench.CONFIGS[name]
n_per = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["n_per"]
niter = int(sys.argv[3]) if len(sys.argv) > 3 else 10
Es, info = bench.make_inputs(cfg, n_per)
m, k, d = cfg["m"], cfg["k"], cfg["d"]
mats = [liger_b200.CountScaled(e, rs, cs) for e, (rs, cs) in zip(Es, info["scales"])] if k <= 32 else Es
ctx = liger_b200.Context(mats, k, 5.0, None, None)
ctx.iterate(niter)
t = time.time()
res = align.quantile_norm_resident(ctx, reference=0)
wall = time.time() - t
H = ctx.download()["H"]
Hn = res["H.norm"]
sizes = [np.bincount(c[c > 0], minlength=k + 1)[1:] for c in res["clusters"]]
changed = float(np.mean(np.abs(Hn - np.vstack(H)).max(axis=1) > 0))
print(json.dumps({"config": name, "cells_per_dataset": n_per, "datasets": d, "k": k, "wall_seconds": wall, **res["info"],
                  "cluster_size_max": int(max(s.max() for s in sizes)), "rows_changed_frac": changed,
                  "finite": bool(np.isfinite(Hn).all())}))
