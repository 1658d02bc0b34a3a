"""Online iNMF at full size (developer tool): C3, minibatch 5000, 5 epochs; prints times and the objective next to a batch run.

    python scripts/online_at_scale.py [config] [n_per] [epochs] [minibatch]
"""
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, liger_b200
from liger_b200 import synthetic as syn
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
cfg = bench.CONFIGS[name]
n_per = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["n_per"]
epochs = int(sys.argv[3]) if len(sys.argv) > 3 else 5
mbs = int(sys.argv[4]) if len(sys.argv) > 4 else 5000
Es, info = bench.make_inputs(cfg, n_per)
m, k, d = cfg["m"], cfg["k"], cfg["d"]
mats = [syn.to_scipy(e) for e in Es]
t = time.time()
out = liger_b200.online_inmf(mats, k=k, lambda_=5.0, maxEpochs=epochs, minibatchSize=mbs, seed=1)
wall = time.time() - t
t = time.time()
batch = liger_b200.inmf(mats, k=k, lambda_=5.0, niter=30, seed=1)
wall_b = time.time() - t
N = n_per * d
print(json.dumps({"config": name, "total_cells": N, "k": k, "epochs": epochs, "minibatch": mbs, "iterations": out["iterations"],
                  "online_loop_ms": out["loop_ms"], "online_wall_seconds": wall, "ms_per_minibatch": out["loop_ms"] / max(out["iterations"], 1),
                  "cells_per_second_loop": N * epochs / (out["loop_ms"] * 1e-3), "online_objErr": out["objErr"],
                  "batch_inmf_30_iterations_objErr": batch["objErr"], "batch_wall_seconds": wall_b,
                  "objective_ratio": out["objErr"] / batch["objErr"]}))
