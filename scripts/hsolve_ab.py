"""Developer probe: total H-solve time over the first N cold-start iterations (C3), for A/B runs of solver knobs."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, liger_b200
cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C3"]
nit = int(sys.argv[2]) if len(sys.argv) > 2 else 20
Es, info = bench.make_inputs(cfg, cfg["n_per"])
m, k, d = cfg["m"], cfg["k"], cfg["d"]
rng = liger_b200.RRandom(1)
W0 = np.abs(rng.runif(m * k, 0, 2)).reshape((m, k), order="F")
V0 = [np.abs(rng.runif(m * k, 0, 2)).reshape((m, k), order="F") for _ in range(d)]
mats = [liger_b200.CountScaled(e, rs, cs) for e, (rs, cs) in zip(Es, info["scales"])]
for rep in range(2):
    ctx = liger_b200.Context(mats, k, 5.0, W0, V0)
    ctx.set_kernel_timing(True)
    ctx.iterate(nit)
    kt = ctx.kernel_times()
    print({a: round(b["ms"] / nit, 4) for a, b in kt.items() if a.startswith("k_nnls_H")}, "loop", round(ctx.timings()["loop_ms"] / nit, 4))
    ctx.close()
