"""Developer probe: how many distinct passive sets do the H columns have?"""
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
prev = None
for it in (1, 2, 5, 10, 20, 30):
    ctx.iterate(it - (0 if prev is None else prev_it))
    prev_it = it
    res = ctx.download(want_passive=True)
    masks = res["H_passive"]
    line = []
    for i in range(min(d, 3)):
        u, c = np.unique(masks[i], return_counts=True)
        c = np.sort(c)[::-1]
        pc = np.array([bin(int(x)).count("1") for x in u[:2000]])
        line.append(f"ds{i}: n={masks[i].size} unique={u.size} top1={c[0]} top10={c[:10].sum()} top100={c[:100].sum()} top1000={c[:1000].sum()}")
    changed = "" if prev is None else f" changed vs prev: {np.mean(masks[0] != prev[0]):.3f}"
    prev = [x.copy() for x in masks]
    print(f"iter {it}: " + " | ".join(line) + changed)
