"""C5: UINMF at scale (2 datasets x 125k cells, 2000 shared + 1000 unshared genes, k=30): runs, monotone objective, timing."""
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import liger_b200
from liger_b200 import synthetic as syn
n = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
Es, Ps, info = syn.make_torch(2, n, 2000, 30, u=1000)
rng = liger_b200.RRandom(1)
m, k, d = 2000, 30, 2
W0 = np.abs(rng.runif(m * k, 0, 2)).reshape((m, k), order="F")
V0 = [np.abs(rng.runif(m * k, 0, 2)).reshape((m, k), order="F") for _ in range(d)]
U0 = [np.abs(rng.runif(1000 * k, 0, 2)).reshape((1000, k), order="F") for _ in range(d)]
# count structure of both blocks (dense-tile tensor-core sweeps over the stacked rows); "gather" as 2nd argument: the fp32 gathers
dense = not (len(sys.argv) > 2 and sys.argv[2] == "gather")
if dense:
    Es = [liger_b200.CountScaled(e, rs, cs) for e, (rs, cs) in zip(Es, info["scales"])]
    Ps = [liger_b200.CountScaled(p, rs, cs) for p, (rs, cs) in zip(Ps, info["unshared_scales"])]
ctx = liger_b200.Context(Es, k, [5.0, 2.0], W0, V0, unsharedList=Ps, Uinit=U0)
ctx.set_kernel_timing(True)
traj = ctx.iterate(10, trajectory=True)
t = ctx.timings()
kt = ctx.kernel_times()
print(json.dumps({"config": "C5", "sweeps": "dense" if dense else "gather", "cells": 2 * n, "nnz": info["nnz"], "ms_per_iter": t["loop_ms"] / 10,
                  "cell_iter_per_s": 2 * n * 10 / (t["loop_ms"] * 1e-3), "monotone": bool(np.all(np.diff(traj) < 0)),
                  "traj_first_last": [traj[0], traj[-1]], "kernels_ms_per_iter": {a: b["ms"] / 10 for a, b in kt.items()}}))
