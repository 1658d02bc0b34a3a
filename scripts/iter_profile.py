"""Developer probe: per-iteration kernel times from a cold start (C3 by default; C4s = one GPU's shard of C4)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, liger_b200
cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C3"]
Es, info = bench.make_inputs(cfg, cfg["n_per"])
m, k, d = cfg["m"], cfg["k"], cfg["d"]
rng = liger_b200.RRandom(1)
W0 = np.abs(rng.runif(m * k, 0, 2)).reshape((m, k), order="F")
V0 = [np.abs(rng.runif(m * k, 0, 2)).reshape((m, k), order="F") for _ in range(d)]
# 2nd argument "gather": plain scaleData (fp32 gather sweeps); default: with its count structure (dense-tile tcgen05 sweeps)
mode = sys.argv[2] if len(sys.argv) > 2 else "dense"
nit = int(sys.argv[3]) if len(sys.argv) > 3 else 30
mats = Es if mode == "gather" else [liger_b200.CountScaled(e, rs, cs) for e, (rs, cs) in zip(Es, info["scales"])]
ctx = liger_b200.Context(mats, k, 5.0, W0, V0)
ctx.set_kernel_timing(True)
for it in range(1, nit + 1):
    ctx.iterate(1)
    t = ctx.timings(); kt = ctx.kernel_times()
    if it <= 8 or it % 5 == 0:
        print(it, round(t["loop_ms"], 3), {a: round(b["ms"], 3) for a, b in kt.items() if b["ms"] > 0.05}, "pivots", t["nnls_pivots"])
