"""Developer check: wall-clock phases of one lgr_inmf call at C3 (LGR_TRACE=1)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, liger_b200, torch
cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C3"]
Es, info = bench.make_inputs(cfg, cfg["n_per"])
m, k, d = cfg["m"], cfg["k"], cfg["d"]
rng = liger_b200.RRandom(1)
W0 = np.abs(rng.runif(m * k, 0, 2)).reshape((m, k), order="F")
V0 = [np.abs(rng.runif(m * k, 0, 2)).reshape((m, k), order="F") for _ in range(d)]
for rep in range(2):
    torch.cuda.synchronize(); t = time.time()
    res = liger_b200.inmf(Es, k=k, lambda_=5.0, niter=30, Winit=W0, Vinit=V0)
    print("inmf wall", time.time() - t, res["timings"])
