"""Developer probe: per-iteration kernel times from a cold start (C3)."""
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
ctx = liger_b200.Context(Es, k, 5.0, W0, V0)
ctx.set_kernel_timing(True)
for it in range(1, 31):
    ctx.iterate(1)
    t = ctx.timings(); kt = ctx.kernel_times()
    if it <= 8 or it % 5 == 0:
        print(it, round(t["loop_ms"], 3), {a: round(b["ms"], 3) for a, b in kt.items() if b["ms"] > 0.05}, "pivots", t["nnls_pivots"])
