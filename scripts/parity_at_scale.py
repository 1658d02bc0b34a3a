"""Full-size parity: CUDA path vs the FP64 C oracle on config C3 (or C3s), same inits, niter iterations."""
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, liger_b200
from liger_b200 import synthetic as syn
from oracle import c_oracle
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
niter = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cfg = bench.CONFIGS[name]
Es, info = bench.make_inputs(cfg, cfg["n_per"])
m, k, d = cfg["m"], cfg["k"], cfg["d"]
rng = liger_b200.RRandom(1)
W0 = np.abs(rng.runif(m * k, 0, 2)).reshape((m, k), order="F")
V0 = [np.abs(rng.runif(m * k, 0, 2)).reshape((m, k), order="F") for _ in range(d)]
# default: scaleData with its count structure -> dense-tile tensor-core sweeps (k <= 32); "gather" as 3rd argument: fp32 gathers
path = sys.argv[3] if len(sys.argv) > 3 else "dense"
mats_in = Es if path == "gather" else [liger_b200.CountScaled(e, rs, cs) for e, (rs, cs) in zip(Es, info["scales"])]
t = time.time()
out = liger_b200.inmf(mats_in, k=k, lambda_=5.0, niter=niter, Winit=W0, Vinit=V0, trajectory=True)
tg = time.time() - t
mats = [syn.to_scipy(e) for e in Es]
t = time.time()
ref = c_oracle.inmf(mats, k, 5.0, niter, W0, V0, trajectory=True)
tc = time.time() - t
rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
res = {"config": name, "niter": niter, "sweeps": path if k <= 32 else (path + (" (two passes: factor halves)" if path == "dense" else "")), "gpu_seconds": tg, "cpu_oracle_seconds": tc, "cpu_threads": c_oracle.num_threads(),
       "rel_W": rel(out["W"], ref["W"]), "rel_V_max": max(rel(a, b) for a, b in zip(out["V"], ref["V"])),
       "rel_H_max": max(rel(a, b) for a, b in zip(out["H"], ref["H"])),
       "obj_traj_max_rel": float(np.max(np.abs(out["traj"] - ref["traj"]) / ref["traj"])),
       "objErr_gpu": out["objErr"], "objErr_oracle": ref["objErr"],
       "H_zero_pattern_mismatch_frac": float(np.mean([np.mean((a == 0) != (b == 0)) for a, b in zip(out["H"], ref["H"])])),
       "H_pattern_mismatch_nondegenerate": int(sum((((a == 0) & (b > 1e-5 * b.max())) | ((b == 0) & (a > 1e-5 * b.max()))).sum()
                                                for a, b in zip(out["H"], ref["H"])))}
print(json.dumps(res))
