"""Resident pipeline at C3 size: counts -> normalize -> scaleNotCenter -> 30 ANLS iterations with X never leaving HBM
(preprocess.Pipeline), against the same steps through HOST buffers (preprocess.normalize / scaleNotCenter, then inmf).
The synthetic matrices of bench.py stand in for the raw counts (any non-negative matrix exercises the same sweeps);
all genes are kept as features so both routes factorise the same scaleData."""
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, liger_b200, torch
from liger_b200 import preprocess as pp, synthetic as syn
cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C3"]
bench.bind_to_gpu_numa(0)
Es, info = bench.make_inputs(cfg, cfg["n_per"])
m, k, d = cfg["m"], cfg["k"], cfg["d"]
raws = [syn.to_scipy(e) for e in Es]          # scipy views of the pinned buffers
genes = [[f"g{j}" for j in range(m)]] * d
numi = [np.ones(cfg["n_per"])] * d
rng = liger_b200.RRandom(1)
W0 = np.abs(rng.runif(m * k, 0, 2)).reshape((m, k), order="F")
V0 = [np.abs(rng.runif(m * k, 0, 2)).reshape((m, k), order="F") for _ in range(d)]
res = {}
for rep in range(4):   # best of 4
    torch.cuda.synchronize(); t0 = time.time()
    pipe = pp.Pipeline(Es, genes, numi)        # the pinned (indptr, indices, data, shape) tuples, in place
    t1 = time.time()
    Xs = pipe.scaleNotCenter(genes[0])
    torch.cuda.synchronize(); t2 = time.time()
    out = liger_b200.inmf(Xs, k=k, lambda_=5.0, niter=30, Winit=W0, Vinit=V0, pinned_outputs=True)
    t3 = time.time()
    if "resident" in res and res["resident"]["total_s"] <= t3 - t0:
        del out, Xs
        pipe.close()
        continue
    res["resident"] = {"upload_counts_s": t1 - t0, "normalize_scale_s": t2 - t1, "inmf_s": t3 - t2, "total_s": t3 - t0,
                       "inmf_timings_ms": out["timings"], "objErr": out["objErr"]}
    Wr = out["W"].copy()
    del out, Xs
    pipe.close()
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    norms = [pp.normalize(r) for r in raws]
    t1 = time.time()
    scaled = [pp.scaleNotCenter(n, genes[0], genes[0]) for n in norms]
    t2 = time.time()
    out = liger_b200.inmf(scaled, k=k, lambda_=5.0, niter=30, Winit=W0, Vinit=V0, pinned_outputs=True)
    t3 = time.time()
    res["host_buffers"] = {"normalize_s": t1 - t0, "scale_s": t2 - t1, "inmf_s": t3 - t2, "total_s": t3 - t0, "objErr": out["objErr"]}
    res["rel_W_between_routes"] = float(np.linalg.norm(out["W"] - Wr) / np.linalg.norm(Wr))
    del out
res["config"] = f"{d} x {cfg['n_per']} cells x {m} genes, nnz {info['nnz']}, k={k}, 30 iterations"
print(json.dumps(res))
