"""Developer probe: a few ANLS iterations at C3 size through the dense-tile sweeps (target of the ncu captures)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import liger_b200  # noqa: E402

cfg = dict(bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C3"])
if len(sys.argv) > 2:
    cfg["n_per"] = int(sys.argv[2])
niter = int(sys.argv[3]) if len(sys.argv) > 3 else 3
Es, info = bench.make_inputs(cfg, cfg["n_per"], pin=False)
m, k, d = cfg["m"], cfg["k"], cfg["d"]
rng = liger_b200.RRandom(1)
W0 = np.abs(rng.runif(m * k, 0, 2)).reshape((m, k), order="F")
V0 = [np.abs(rng.runif(m * k, 0, 2)).reshape((m, k), order="F") for _ in range(d)]
mats = [liger_b200.CountScaled(e, rs, cs) for e, (rs, cs) in zip(Es, info["scales"])]
if os.environ.get("NO_DENSE"):
    mats = Es
ctx = liger_b200.Context(mats, k, 5.0, W0, V0)
ctx.set_kernel_timing(True)
ctx.iterate(niter)
kt = ctx.kernel_times()
print({a: round(b["ms"] / max(b["launches"], 1), 4) for a, b in kt.items()}, "obj", ctx.objective())
