#!/usr/bin/env python
"""bench.py -- iNMF cell*iterations/s on synthetic Poisson-sparse counts (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C3|C4|tiny]
                    [--scaling strong|weak]

A "step" is one ANLS iteration (H_i for all i -> V_i -> W) over ALL cells of the workload.
  value : whole-job cell*iterations/s with inputs resident in HBM: K iterations FROM A COLD START (the seeded inits,
          empty passive sets) plus the final objective, as SURVEY 8d defines the metric (lgr_ctx_run; CUDA events on the
          library's stream; max over ranks).  The W warm-up iterations run first and the context is then reset.
  e2e   : the same metric through the reference-facing call lgr_inmf(niter = K) with HOST buffers (H2D of X / inits and
          D2H of W, V, H inside the timed region): page-locked buffers (the contract's figure) and, under
          e2e.pageable, ordinary pageable memory as an R caller would hand over
  roofline     : dominant kernel, algorithmic bytes per launch / its mean CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline : oracle/inmf_oracle.c (FP64 + OpenMP, "CPU restatement of RcppPlanc inmf"): the whole workload on all
                 host cores, and a 1/8 sample at the reference's default nCores = 2 (R/integration.R:240)
Under torchrun (N > 1) the scaling is STRONG by default: the configuration's cells (C3: 1 M) are split over the N ranks,
every rank holding a cell shard of every dataset (BASELINE.json: "1M cells ... 1/2/4/8 B200").  --scaling weak keeps the
cells per GPU fixed instead; at N > 1 the weak figure is also reported under "extra".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: d, cells per dataset (per GPU), genes, k, density
    "C3": dict(d=8, n_per=125000, m=2000, k=30, density=0.05),     # BASELINE.json configs[2], the metric's config
    "C3s": dict(d=8, n_per=12500, m=2000, k=30, density=0.05),     # 1/10 sample (cpu legs, smoke runs)
    "C4": dict(d=16, n_per=250000, m=3000, k=50, density=0.05),    # configs[3] (8-GPU shape; per-GPU shard = 1/8)
    "C4s": dict(d=16, n_per=31250, m=3000, k=50, density=0.05),    # one GPU's 1/8 cell shard of C4
    "tiny": dict(d=2, n_per=3000, m=500, k=20, density=0.05),
}
LAMBDA = 5.0
NO_DENSE = False


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        time.sleep(0.15)
        self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.samples:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:  # region shorter than the sampling period: take everything we have
            for ts, line in self.samples:
                parts = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(parts[0]))
                    mx.append(float(parts[1]))
                except Exception:
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa(index):
    """Pin this process to the CPU cores local to GPU `index` (sysfs local_cpulist of its PCI function), so that the
    page-locked host buffers it allocates sit on the NUMA node the GPU hangs off: a remote node halves the H2D rate
    (measured: 25 ms vs 45 ms for the 1.2 GB of C3).  One process per GPU, each bound to its own GPU's node."""
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(index)],
                             capture_output=True, text=True, timeout=10).stdout.strip().splitlines()[0].strip().lower()
        if out.startswith("00000000:"):
            out = "0000:" + out[9:]
        cl = open(f"/sys/bus/pci/devices/{out}/local_cpulist").read().strip()
        cpus = set()
        for part in cl.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception:
        pass
    return None


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed ncu --set full capture
    of this same command (profiles/r*_traffic.json, written by scripts/make_profiles.py); None if not captured."""
    import glob
    alias = {"k_nnls_H": "k_nnls_reg", "k_spmm_ltx": "k_spmm_ltx_tiled", "k_spmm_xh": "k_spmm_xh",
             "k_dense_ltx": "k_dense_ltx", "k_dense_xh": "k_dense_xh", "k_gram_tc": "k_gram_tc"}
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))
    if not files or kernel not in alias:
        return None
    try:
        t = json.load(open(files[-1]))
        for name, v in t.items():
            if name.startswith(alias[kernel]):
                return float(v)
    except Exception:
        pass
    return None


def algorithmic_bytes(cfg, nnz_total, n_total):
    """Algorithmic bytes of ONE iteration (DESIGN.md 'roofline', SURVEY 8d convention): two sweeps over X (fp32 value +
    int32 index + column pointers), H written once and read once, B written+read, W/V read+written."""
    d, m, k = cfg["d"], cfg["m"], cfg["k"]
    x_sweep = 8 * nnz_total + 4 * (n_total + d)
    return {"ltx": x_sweep + 4 * n_total * k,                              # read X, write B (L is cache resident)
            "nnls_H": 2 * 4 * n_total * k + 16 * n_total,                  # read B, write H, masks in/out
            "xh": x_sweep + 4 * n_total * k,                               # read X, read H
            "iteration": 2 * x_sweep + 12 * n_total * k + 8 * m * k * (d + 1)}   # SURVEY 8d formula


def kernel_class(name):
    """Which algorithmic-byte formula a kernel of the loop falls under (None: a small dense kernel, no HBM roofline)."""
    if "ltx" in name:
        return "ltx"
    if "xh" in name:
        return "xh"
    if name == "k_nnls_H":
        return "nnls_H"
    return None


def host_threads():
    """Threads the CPU legs use: every core this process may run on (never the OpenMP default -- torchrun exports
    OMP_NUM_THREADS=1)."""
    try:
        return max(len(os.sched_getaffinity(0)), 1)
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, cfg, rank):
    """--impl reference: the CPU restatement (oracle/inmf_oracle.c) on all host cores, on the WHOLE workload of the
    configuration (C3: 8 x 125 000 cells), K iterations from the same cold start."""
    if rank != 0:
        return
    from oracle import c_oracle, inmf_oracle as orc
    from liger_b200 import synthetic as syn
    frac = args.cpu_fraction
    n_s = max(int(cfg["n_per"] * frac), cfg["k"] + 1)
    Es, _ = make_inputs(cfg, n_s, device=None if have_cuda() else "cpu", pin=False)
    mats = [syn.to_scipy(e) for e in Es]
    W0, V0, _, _ = orc.seeded_init(1, cfg["m"], cfg["k"], [n_s] * cfg["d"])
    nth = host_threads()
    N = n_s * cfg["d"]
    for _ in range(min(args.warmup, 1)):
        c_oracle.inmf(mats, cfg["k"], LAMBDA, 1, W0, V0, nthreads=nth)
    t0 = time.time()
    out = c_oracle.inmf(mats, cfg["k"], LAMBDA, args.steps, W0, V0, nthreads=nth)
    dt = time.time() - t0
    val = N * args.steps / dt
    sample = (f"{cfg['d']} datasets x {n_s} cells ({frac:.3g} of the configuration's cells), {args.steps} iterations from the "
              f"seeded cold start, FP64, OpenMP x{nth}, {dt:.1f} s")
    line = {"impl": "reference", "metric": "iNMF cell*iterations/s", "value": val, "unit": "cell*iter/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(cfg, args, max(args.gpus, 1), "strong", n_s * cfg["d"],
                                      "datasets" if (args.gpus > 1 and args.placement != "cells" and cfg["d"] % args.gpus == 0 and
                                                     args.scaling == "strong") else "cells"),
            "cpu_baseline": {"value": val, "unit": "cell*iter/s", "cores": nth, "kind": "port", "sample": sample,
                             "note": "CPU restatement of RcppPlanc inmf (oracle/inmf_oracle.c); the reference's R/RcppPlanc stack cannot be built here"},
            "e2e": {"value": val, "unit": "cell*iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "objErr": out["objErr"]}
    print(json.dumps(line))


def have_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def make_inputs(cfg, n_per, device=None, value_dtype="float64", shard=0, pin=True):
    from liger_b200 import synthetic as syn
    Es, Ps, info = syn.make_torch(cfg["d"], n_per, cfg["m"], cfg["k"], density=cfg["density"], device=device,
                                  value_dtype=value_dtype, shard=shard, pin=pin)
    return Es, info


def workload_config(cfg, args, world, scaling, total_cells, placement="cells"):
    per = ("in total, whole datasets placed on the GPUs" if placement == "datasets" and world > 1 else
           "in total, split by cells over the GPUs") if scaling == "strong" else "per GPU"
    return {"workload": f"{args.config}: synthetic planted-Poisson iNMF, {cfg['d']} datasets x {cfg['n_per']} cells x "
                        f"{cfg['m']} genes {per}, {int(cfg['density'] * 100)}% density, k={cfg['k']}, lambda={LAMBDA}, "
                        f"{args.steps} ANLS iterations from the seeded cold start + final objective",
            "datasets": cfg["d"], "cells_per_dataset": cfg["n_per"], "genes": cfg["m"], "k": cfg["k"],
            "density": cfg["density"], "total_cells": int(total_cells),
            "parallelism": (f"dataset-shard x{world} ({cfg['d'] // world} whole dataset(s) per GPU)" if placement == "datasets" and world > 1
                            else f"cell-shard x{world}"),
            "scaling": scaling,
            "l2_policy": "inputs larger than L2 at N <= 4 (X is >= 0.35 GB per GPU in two layouts); every iteration also "
                         "rewrites B, H and the statistics, so no sweep finds its input in L2 from the previous one"}


class Job:
    """One resident workload on this rank: inputs, inits, context."""

    def __init__(self, cfg, n_per_rank, rank, world, local_rank, nccl_id, pin=True, placement="cells"):
        import liger_b200
        self.rank, self.world, self.placement = rank, world, placement
        d_all, m, k = cfg["d"], cfg["m"], cfg["k"]
        if placement == "datasets":   # whole-dataset placement: this rank holds d / world datasets with ALL their cells
            d = d_all // world
            cfg = dict(cfg, d=d)
            first = rank * d
        else:                         # cell sharding: this rank holds n_per_rank cells of every dataset
            d, first = d_all, 0
        self.cfg = cfg                # the rank-local view (algorithmic bytes, tensor flops)
        t = time.time()
        self.Es, self.info = make_inputs(cfg, n_per_rank, value_dtype="float64", shard=rank, pin=pin)
        self.gen_seconds = time.time() - t
        self.n_local = d * n_per_rank
        self.n_total = self.n_local * world
        rng = liger_b200.RRandom(1)       # R's set.seed(1) stream, the draw order of .runINMF.list
        self.W0 = np.abs(rng.runif(m * k, 0, 2)).reshape((m, k), order="F")
        V_all = [np.abs(rng.runif(m * k, 0, 2)).reshape((m, k), order="F") for _ in range(d_all)]
        self.V0 = V_all[first:first + d]
        # scaleData with its count structure (values = count * 1/rms[gene] * 1/library size[cell]): the library verifies it
        # on upload and runs the two sweeps over X on the tensor cores; --no-dense runs the fp32 gather sweeps instead
        self.mats = self.Es if NO_DENSE else [liger_b200.CountScaled(e, rs, cs) for e, (rs, cs) in zip(self.Es, self.info["scales"])]
        self.ctx = liger_b200.Context(self.mats, k, LAMBDA, self.W0, self.V0, device=local_rank, rank=rank, world=world,
                                      nccl_id=nccl_id, flags=(0x80 if placement == "datasets" else 0x8) if world > 1 else 0)

    def close(self):
        self.ctx.close()


def timed_runs(job, args, barrier, allmax, sampler=None):
    """warm-up, then K iterations from the cold start + final objective (un-instrumented: `value`), then the same K steps
    again with CUDA events around every kernel (roofline), then K more from the converged state (warm figure)."""
    ctx = job.ctx
    ctx.set_kernel_timing(False)
    ctx.iterate(args.warmup)
    ctx.reset(job.W0, job.V0)
    if sampler:
        sampler.start()
        time.sleep(0.25)
    barrier()
    t0 = time.time()
    obj = ctx.run(args.steps)
    barrier()
    t1 = time.time()
    tm = ctx.timings()
    clocks = sampler.stop(t0, t1) if sampler else None
    loop_ms = allmax(tm["loop_ms"])
    launches = int(tm["kernel_launches"])
    # instrumented pass over the same region (same inits, same K): per-kernel CUDA events on the launch stream
    ctx.reset(job.W0, job.V0)
    ctx.set_kernel_timing(True)
    barrier()
    ctx.run(args.steps)
    tmi = ctx.timings()
    kt = ctx.kernel_times()
    ctx.set_kernel_timing(False)
    barrier()
    ctx.iterate(args.steps)     # warm: passive sets and factors of a converged run
    warm_ms = allmax(ctx.timings()["loop_ms"])
    return {"obj": obj, "loop_ms": loop_ms, "launches": launches, "clocks": clocks, "kt": kt,
            "instrumented_loop_ms": allmax(tmi["loop_ms"]), "h_phase_ms": tmi["h_phase_ms"], "vw_phase_ms": tmi["vw_phase_ms"],
            "comm_ms": allmax(tmi["comm_ms"]), "warm_ms": warm_ms, "pivots": tm["nnls_pivots"],
            "unconverged": tm["nnls_unconverged"], "wall_s": t1 - t0}


def roofline_of(job, res, args, hbm_peak, peak_src, with_traffic):
    """`roofline` of the bench line: the kernel with the largest share of the step (by its CUDA-event time) against the HBM
    roofline, with the same figure for every large kernel under `per_kernel` and the whole iteration under `iteration`."""
    ab = algorithmic_bytes(job.cfg, job.info["nnz"], job.n_local)
    kt = res["kt"]
    cands = {kn: v for kn, v in kt.items() if kernel_class(kn)}
    if not cands:
        return None
    dom = max(cands.items(), key=lambda kv: kv[1]["ms"])[0]
    per_launch_ms = kt[dom]["ms"] / max(kt[dom]["launches"], 1)
    bytes_launch = ab[kernel_class(dom)]
    ach = bytes_launch / (per_launch_ms * 1e-3) / 1e9
    it_ms = res["loop_ms"] / args.steps
    notes = {"nnls_H": "the H solve moves 2 x 4 n k + 16 n bytes but is bound by the SM's shared-memory data path and by latency "
                       "(232 registers per thread, 8 warps per SM; ncu: l1tex 68 %, dram 4 %), not by HBM: see DESIGN.md section 5",
             "ltx": "dense-tile tcgen05 sweep: bound by the tensor pipe / operand staging, see roofline.tensor",
             "xh": "dense-tile tcgen05 sweep: bound by the tensor pipe / operand staging, see roofline.tensor"}
    # tensor-pipe view of the two dense sweeps: algorithmic MMA flops (M = 128 rows incl. padding, N = cells or genes, K) over
    # the measured dense bf16 peak of MEASURED_PEAKS.json
    tensor = None
    if any(kn.startswith("k_dense") for kn in kt):
        try:
            pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            tf_peak = float(pk.get("bf16_tflops_sustained", pk.get("bf16_tflops", 1590.0)))
        except Exception:
            tf_peak = 1400.0
        cfg = job.cfg
        gp = ((cfg["m"] + 63) // 64) * 64
        halves = 2 if cfg["k"] > 32 else 1        # k > 32: one pass per half of the factors
        flops = 2.0 * 128 * gp * job.n_local * halves   # per sweep: every (cell, gene) pair meets the 128-row operand slab once per pass
        tensor = {"peak_tflops": tf_peak, "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (cuBLAS bf16 GEMM, seconds-long loop)",
                  "note": "flops of the issued tcgen05.mma tiles (128 x 256 x 16 bf16, 96 of the 128 rows carry data)"}
        for kn in ("k_dense_ltx", "k_dense_xh"):
            if kn in kt:
                ms = kt[kn]["ms"] / max(kt[kn]["launches"], 1)
                tensor[kn] = {"tflops": flops / (ms * 1e-3) / 1e12, "frac": flops / (ms * 1e-3) / 1e12 / tf_peak}
    return {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
            "traffic": ncu_traffic(dom) if with_traffic else None, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": bytes_launch, "ms_per_launch": per_launch_ms, "note": notes.get(kernel_class(dom)),
            "timing": "CUDA events on the library's launch stream around every launch, over the same K cold-start steps "
                      "run a second time (the `value` pass itself carries no per-kernel events)",
            "iteration": {"algorithmic_bytes": ab["iteration"], "achieved": ab["iteration"] / (it_ms * 1e-3) / 1e9,
                          "frac": ab["iteration"] / (it_ms * 1e-3) / 1e9 / hbm_peak},
            "per_kernel": {kn: {"ms_per_iter": v["ms"] / args.steps,
                                "frac_of_hbm": (ab[kernel_class(kn)] / (v["ms"] / max(v["launches"], 1) * 1e-3) / 1e9 / hbm_peak)
                                if kernel_class(kn) else None,
                                "traffic": ncu_traffic(kn) if with_traffic else None} for kn, v in kt.items()},
            "tensor": tensor,
            "instrumented_ms_per_step": res["instrumented_loop_ms"] / args.steps,
            "phases_ms_per_iter": {"h_phase": res["h_phase_ms"] / args.steps, "vw_phase": res["vw_phase_ms"] / args.steps,
                                   "comm": res["comm_ms"] / args.steps}}


def sharded_parity(rank, world, local_rank, nccl_id, dist, torch):
    """Small fixed-seed run THROUGH THE SAME SHARDED PATH (cell shards + the per-iteration all-reduce), against the
    reference's own golden factors of runINMF(pbmc, k = 20, lambda = 5, nIteration = 30, seed = 1) (data/pbmcPlot.rda,
    stored in tests/golden/pbmc_golden.npz): relative Frobenius errors and whether W, V are identical on every rank."""
    import scipy.sparse as sp
    import liger_b200
    path = os.path.join(ROOT, "tests", "golden", "pbmc_golden.npz")
    if not os.path.exists(path):
        return None
    g = np.load(path)
    names = ("ctrl", "stim")
    Es = [sp.csc_matrix((g[f"scale_{n}_data"], g[f"scale_{n}_indices"], g[f"scale_{n}_indptr"]),
                        shape=tuple(g[f"scale_{n}_shape"])) for n in names]
    m, k = Es[0].shape[0], 20
    rng = liger_b200.RRandom(1)
    W0 = np.abs(rng.runif(m * k, 0, 2)).reshape((m, k), order="F")
    V0 = [np.abs(rng.runif(m * k, 0, 2)).reshape((m, k), order="F") for _ in Es]
    out = liger_b200.inmf(Es, k=k, lambda_=5.0, niter=30, Winit=W0, Vinit=V0, device=local_rank, rank=rank, world=world,
                          nccl_id=nccl_id)      # slice mode: the library takes this rank's columns
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
    rows = g["golden_rows"]
    errs = [rel(out["W"][rows], g["golden_W"])] + [rel(out["V"][i][rows], g[f"golden_V_{n}"]) for i, n in enumerate(names)]
    herr = 0.0
    for i, n in enumerate(names):
        nf = Es[i].shape[1]
        a, b = nf * rank // world, nf * (rank + 1) // world
        herr = max(herr, rel(out["H"][i][a:b].T, g[f"golden_H_{n}"][:, a:b]))
    identical = True
    if world > 1:
        t = torch.tensor(np.concatenate([out["W"].ravel()] + [v.ravel() for v in out["V"]]), device="cuda")
        t0 = t.clone()
        dist.broadcast(t0, 0)
        diff = (t - t0).abs().max().reshape(1)
        dist.all_reduce(diff, op=dist.ReduceOp.MAX)
        identical = bool(float(diff.item()) == 0.0)
        e = torch.tensor([max(errs), herr], dtype=torch.float64, device="cuda")
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
        werr, herr = float(e[0].item()), float(e[1].item())
    else:
        werr = max(errs)
    return {"case": "pbmc k=20 lambda=5 30 iterations seed=1 vs the reference's golden factors (data/pbmcPlot.rda), run "
                    f"through the {world}-rank cell-sharded path", "W_V_rel_err": werr, "H_rel_err": herr,
            "objErr": out["objErr"], "objErr_reference_value": 36105.069129349,
            "ranks_identical_W_V": identical, "pass": bool(werr < 1e-3 and herr < 1e-3 and identical)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="C3")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--cpu-fraction", type=float, default=1.0, help="cells fraction for the CPU legs")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the weak-scaling / C4 extras at N > 1")
    ap.add_argument("--placement", choices=["auto", "cells", "datasets"], default="auto",
                    help="multi-GPU placement of the strong-scaling job: whole datasets per rank when the dataset count divides "
                         "(auto), or cell shards of every dataset")
    ap.add_argument("--no-dense", action="store_true", help="pass scaleData without its count structure: fp32 gather sweeps")
    args = ap.parse_args()
    global NO_DENSE
    NO_DENSE = args.no_dense
    if args.impl != "reference":
        args.warmup = max(args.warmup, 3)
    cfg = dict(CONFIGS[args.config])
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return

    import torch
    import torch.distributed as dist
    import liger_b200
    from liger_b200 import synthetic as syn
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the iNMF path has no CPU fallback; use --impl reference for the CPU leg)")
    cpus = bind_to_gpu_numa(local_rank)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    hbm_peak, peak_src = peaks()

    nccl_id = None
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(liger_b200.nccl_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().numpy().tobytes())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- the workload: strong scaling splits the configuration's cells over the ranks; every rank generates ITS shard ----
    m, k = cfg["m"], cfg["k"]
    # strong scaling with d divisible by N: whole-dataset placement (V_i / U_i solves local, one small all-reduce for W);
    # otherwise cell sharding (every rank holds a slice of every dataset, statistics all-reduced)
    placement = "cells"
    if world > 1 and args.scaling == "strong" and args.placement != "cells" and cfg["d"] % world == 0:
        placement = "datasets"
    if args.placement == "datasets" and placement != "datasets":
        raise SystemExit("--placement datasets needs --scaling strong and a dataset count divisible by the number of GPUs")
    n_rank = cfg["n_per"] if placement == "datasets" else (cfg["n_per"] // world if args.scaling == "strong" else cfg["n_per"])
    job = Job(cfg, n_rank, rank, world, local_rank, nccl_id, placement=placement)
    d = job.cfg["d"]      # datasets on this rank
    sampler = ClockSampler(local_rank)
    res = timed_runs(job, args, barrier, allmax, sampler)
    value = job.n_total * args.steps / (res["loop_ms"] * 1e-3)
    roofline = roofline_of(job, res, args, hbm_peak, peak_src, args.config == "C3" and world == 1)
    if world > 1 or args.no_extra:
        job.close()      # (at N = 1 the context lives on for the quantileNorm extra, which runs after the e2e measurements)

    # ---- e2e: lgr_inmf with HOST buffers (upload + K iterations from a cold start + objective + download), through the
    #      public API at N GPUs: every rank passes its own host shard; timed barrier-to-barrier, max over ranks;
    #      median of 3 calls (the PCIe copy rate of a fresh box varies from call to call) ----
    e2e = None
    Es, W0, V0 = job.Es, job.W0, job.V0
    scales = job.info["scales"]
    wrap = (lambda mats: mats) if NO_DENSE else (lambda mats: [liger_b200.CountScaled(e, rs, cs) for e, (rs, cs) in zip(mats, scales)])
    if not args.no_e2e:
        from liger_b200.api import pinned_empty
        n_loc = [int(e[3][1]) for e in Es]
        # result buffers of the page-locked series: allocated once, like the inputs (a host that keeps pinned buffers)
        pinned_out = {"W": pinned_empty((m, k)), "V": [pinned_empty((m, k)) for _ in range(d)],
                      "H": [pinned_empty((n, k)) for n in n_loc]}

        def call(mats, pinned):
            return liger_b200.inmf(wrap(mats), k=k, lambda_=LAMBDA, niter=args.steps, Winit=W0, Vinit=V0,
                                   outputs=pinned_out if pinned else None,
                                   pinned_io=pinned, device=local_rank, rank=rank, world=world, nccl_id=nccl_id,
                                   local_shard=placement == "cells", dataset_shard=placement == "datasets")
            # (same unique id => the library reuses its communicator)

        def series(mats, pinned, reps):
            runs = []
            for _rep in range(reps):
                barrier()
                t0e = time.time()
                r = call(mats, pinned)
                barrier()
                runs.append((allmax(time.time() - t0e), r["timings"], r["objErr"],
                             sum(h.nbytes for h in r["H"]) + W0.nbytes * (d + 1)))
                del r
            runs.sort(key=lambda x: x[0])
            return runs
        _w = call(Es, True)   # warm the memory pools
        del _w
        runs = series(Es, True, 3)
        dt, tms, obj_e, d2h = runs[len(runs) // 2]
        h2d = sum(e[0].nbytes + e[1].nbytes + e[2].nbytes for e in Es) + W0.nbytes * (d + 1)
        if not NO_DENSE:
            h2d += sum(rs.nbytes + cs.nbytes for rs, cs in scales)
        e2e = {"value": job.n_total * args.steps / dt, "unit": "cell*iter/s", "h2d_bytes_per_step": h2d * world / args.steps,
               "d2h_bytes_per_step": d2h * world / args.steps, "seconds": dt, "seconds_all_calls": [r[0] for r in runs],
               "timings_ms": tms, "objErr": obj_e, "host_memory": "page-locked",
               "note": "one lgr_inmf call per GPU = H2D of that rank's X shard (dgCMatrix slots, f64 values) + inits, "
                       "device layout build, K ANLS iterations from a cold start, objective, D2H of W/V/H (f64); "
                       "barrier-to-barrier, max over ranks, median of 3 calls"}
        # the same call on PAGEABLE buffers (what an Rcpp caller owns): inputs copied into ordinary numpy arrays, outputs
        # freshly allocated by numpy; the library stages them through its page-locked bounce ring
        Ep = [(np.array(e[0]), np.array(e[1]), np.array(e[2]), e[3]) for e in Es]
        runs_p = series(Ep, False, 2)
        dtp, tmsp, _, _ = runs_p[0]
        e2e["pageable"] = {"value": job.n_total * args.steps / dtp, "unit": "cell*iter/s", "seconds": dtp,
                           "seconds_all_calls": [r[0] for r in runs_p], "timings_ms": tmsp, "host_memory": "pageable (numpy)",
                           "note": "best of 2 calls"}
        del Ep

    parity = sharded_parity(rank, world, local_rank, nccl_id, dist, torch)

    # ---- extras at N > 1: the weak-scaling figure (cells per GPU fixed) and, at N = 8, C4 (4 M cells x 3000 genes x 16
    #      datasets, k = 50) sharded over the GPUs ----
    extra = {"warm_ms_per_step": res["warm_ms"] / args.steps,
             "warm_value": job.n_total * args.steps / (res["warm_ms"] * 1e-3),
             "warm_note": "K further iterations from the converged state of the timed run (warm passive sets)"}
    # the step after the path, on the H that is still in HBM (single GPU): quantileNorm with the reference's defaults
    align_extra = None
    if world == 1 and not args.no_extra:
        try:
            from liger_b200 import align as _align
            t0a = time.time()
            qn = _align.quantile_norm_resident(job.ctx, reference=0)
            align_extra = {"what": "quantileNorm (reference defaults: 50 quantiles, 20 approximate neighbours eps = 0.9, in-place vote) of "
                                   "the resident H of the timed run (measured after everything else: its host threads and allocations must not disturb the e2e calls)",
                           "seconds": time.time() - t0a, **qn["info"]}
            del qn
        except Exception as e:      # never let the extra take the bench line down
            align_extra = {"error": str(e)[:200]}
    if align_extra is not None:
        extra["quantile_norm"] = align_extra
    if world == 1 and not args.no_extra:
        job.close()
    del Es
    job.Es = None
    if world > 1 and not args.no_extra:
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        other = "weak" if args.scaling == "strong" else "strong"
        n2 = cfg["n_per"] if other == "weak" else cfg["n_per"] // world
        j2 = Job(cfg, n2, rank, world, local_rank, nccl_id, pin=False)
        r2 = timed_runs(j2, args, barrier, allmax)
        extra[other] = {"scaling": other, "total_cells": j2.n_total, "value": j2.n_total * args.steps / (r2["loop_ms"] * 1e-3),
                        "ms_per_step": r2["loop_ms"] / args.steps, "comm_ms_per_step": r2["comm_ms"] / args.steps,
                        "objErr": r2["obj"]}
        j2.close()
        del j2
        if world == 8 and args.config == "C3":
            gc.collect()
            torch.cuda.empty_cache()
            c4 = dict(CONFIGS["C4"])
            # same placement rule as the headline run: 16 datasets over 8 ranks = two whole datasets per GPU, unless cell
            # sharding was asked for
            p4 = "datasets" if args.placement != "cells" else "cells"
            j4 = Job(c4, c4["n_per"] if p4 == "datasets" else c4["n_per"] // world, rank, world, local_rank, nccl_id, pin=False,
                     placement=p4)
            r4 = timed_runs(j4, args, barrier, allmax)
            rf4 = roofline_of(j4, r4, args, hbm_peak, peak_src, False)
            extra["C4"] = {"workload": "C4: 16 datasets x 250000 cells x 3000 genes (4 M cells) over 8 GPUs (" +
                                       ("two whole datasets per GPU" if p4 == "datasets" else "cell shards") + "), 5% density, k=50",
                           "total_cells": j4.n_total, "value": j4.n_total * args.steps / (r4["loop_ms"] * 1e-3),
                           "ms_per_step": r4["loop_ms"] / args.steps, "warm_ms_per_step": r4["warm_ms"] / args.steps,
                           "comm_ms_per_step": r4["comm_ms"] / args.steps, "objErr": r4["obj"],
                           "roofline_iteration_frac": rf4["iteration"]["frac"] if rf4 else None,
                           "kernels_ms_per_iter": {kn: v["ms_per_iter"] for kn, v in rf4["per_kernel"].items()} if rf4 else None}
            j4.close()
            del j4

    # ---- CPU baseline (rank 0, N = 1 only): the whole workload on all host cores, and a 1/8 sample at nCores = 2 ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import c_oracle
        Esc, _ = make_inputs(cfg, cfg["n_per"], value_dtype="float64", shard=rank, pin=False)
        frac = args.cpu_fraction
        n_s = max(int(cfg["n_per"] * frac), k + 1)
        cut = lambda e, n: syn.to_scipy((e[0][:n + 1], e[1][:e[0][n]], e[2][:e[0][n]], (m, n)))
        mats = [cut(e, n_s) for e in Esc]
        nth = host_threads()
        it_cpu = args.steps
        tc0 = time.time()
        c_oracle.inmf(mats, k, LAMBDA, it_cpu, W0, V0, nthreads=nth)
        dtc = time.time() - tc0
        cpu = {"value": d * n_s * it_cpu / dtc, "unit": "cell*iter/s", "cores": nth, "kind": "port",
               "sample": f"first {n_s} cells of each of the {d} datasets ({frac:.3g} of the workload), {it_cpu} iterations from "
                         f"the same init, FP64, OpenMP x{nth} ({dtc:.1f} s); CPU restatement of RcppPlanc inmf "
                         f"(oracle/inmf_oracle.c)"}
        n2 = max(cfg["n_per"] // 8, k + 1)
        mats2 = [cut(e, n2) for e in Esc]
        tc0 = time.time()
        c_oracle.inmf(mats2, k, LAMBDA, it_cpu, W0, V0, nthreads=2)
        dt2 = time.time() - tc0
        cpu["ncores2"] = {"value": d * n2 * it_cpu / dt2, "unit": "cell*iter/s", "cores": 2,
                          "sample": f"first {n2} cells of each dataset (1/8 of the workload), {it_cpu} iterations, OpenMP x2 "
                                    f"({dt2:.1f} s): the reference's default nCores = 2L (R/integration.R:240)"}

    if rank == 0:
        line = {"metric": "iNMF cell*iterations/s", "value": value, "unit": "cell*iter/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["loop_ms"] / args.steps, "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "f32 (fp64 Gram / V,W solves / objective)",
                "data": "synthetic", "config": workload_config(cfg, args, world, args.scaling, job.n_total, placement),
                "clocks": res["clocks"], "e2e": e2e, "gpu_launches": res["launches"], "roofline": roofline,
                "cpu_baseline": cpu, "parity": parity, "extra": extra,
                "objErr": res["obj"], "nnz_per_gpu": job.info["nnz"], "gen_seconds": job.gen_seconds,
                "host_cpus_bound": (f"{len(cpus)} cores local to the GPU" if cpus else None),
                "nnls": {"pivots": res["pivots"], "unconverged": res["unconverged"]}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
