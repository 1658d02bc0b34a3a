"""Developer probe: histogram of final passive-set sizes of the H solves (old kernel's hist mode)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, liger_b200
cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C3s"]
Es, info = bench.make_inputs(cfg, cfg["n_per"])
m, k, d = cfg["m"], cfg["k"], cfg["d"]
rng = liger_b200.RRandom(1)
W0 = np.abs(rng.runif(m * k, 0, 2)).reshape((m, k), order="F")
V0 = [np.abs(rng.runif(m * k, 0, 2)).reshape((m, k), order="F") for _ in range(d)]
ctx = liger_b200.Context(Es, k, 5.0, W0, V0)
for it in (1, 2, 3, 5, 10, 20, 30):
    ctx.iterate(1 if it <= 3 else it - prev)
    prev = it
    res = ctx.download(want_passive=True)
    pc = np.concatenate([np.array([bin(int(x)).count("1") for x in mm]) for mm in res["H_passive"][:2]])
    print(it, "p histogram:", np.bincount(pc, minlength=k + 1).tolist())
