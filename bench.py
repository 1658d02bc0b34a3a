#!/usr/bin/env python
"""bench.py -- iNMF cell*iterations/s on synthetic Poisson-sparse counts (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C3|C3s|C4|tiny]

A "step" is one ANLS iteration (H_i for all i -> V_i -> W) over ALL cells of the workload.
  value : whole-job cell*iterations/s with inputs resident in HBM (lgr_ctx_iterate, CUDA events, max over ranks)
  e2e   : the same metric through the reference-facing call lgr_inmf(niter = K) with HOST buffers
          (H2D of X / inits and D2H of W, V, H inside the timed region)
  roofline     : dominant kernel, algorithmic bytes per launch / its mean CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline : oracle/inmf_oracle.c (FP64 + OpenMP, "CPU restatement of RcppPlanc inmf") on a bounded sample
Under torchrun (N > 1) every rank holds a cell shard of every dataset (weak scaling: cells per GPU fixed).
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
    alias = {"k_nnls_H": "k_nnls_reg", "k_spmm_ltx": "k_spmm_ltx_tiled", "k_spmm_xh": "k_spmm_xh"}
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
    """Algorithmic bytes of ONE iteration (DESIGN.md 'roofline'): two sweeps over X (fp32 value + int32 index
    + column pointers), H written once and read once, B written+read, W/V read+written."""
    d, m, k = cfg["d"], cfg["m"], cfg["k"]
    x_sweep = 8 * nnz_total + 4 * (n_total + d)
    return {"k_spmm_ltx": x_sweep + 4 * n_total * k,                       # read X, write B (L is cache resident)
            "k_nnls_H": 2 * 4 * n_total * k + 16 * n_total,                # read B, write H, masks in/out
            "k_spmm_xh": x_sweep + 4 * n_total * k,                        # read X (tiled CSR), read H
            "iteration": 2 * x_sweep + 12 * n_total * k + 8 * m * k * (d + 1)}   # SURVEY 8d formula


def run_reference(args, cfg, rank):
    """--impl reference: the CPU restatement (oracle/inmf_oracle.c, all host threads) on a bounded sample."""
    if rank != 0:
        return
    from oracle import c_oracle, inmf_oracle as orc
    from liger_b200 import synthetic as syn
    frac = args.cpu_fraction
    n_s = max(int(cfg["n_per"] * frac), cfg["k"] + 1)
    Es, _ = make_inputs(cfg, n_s, device="cpu" if not have_cuda() else None)
    mats = [syn.to_scipy(e) for e in Es]
    W0, V0, _, _ = orc.seeded_init(1, cfg["m"], cfg["k"], [n_s] * cfg["d"])
    nth = c_oracle.num_threads()
    N = n_s * cfg["d"]
    for _ in range(min(args.warmup, 1)):
        c_oracle.inmf(mats, cfg["k"], LAMBDA, 1, W0, V0, nthreads=nth)
    t0 = time.time()
    out = c_oracle.inmf(mats, cfg["k"], LAMBDA, args.steps, W0, V0, nthreads=nth)
    dt = time.time() - t0
    val = N * args.steps / dt
    sample = f"{cfg['d']} datasets x {n_s} cells ({frac:.3g} of the workload's cells per dataset), {args.steps} iterations, FP64, OpenMP"
    line = {"impl": "reference", "metric": "iNMF cell*iterations/s", "value": val, "unit": "cell*iter/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 / args.steps * (cfg["n_per"] / n_s),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(cfg, args, 1),
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


def make_inputs(cfg, n_per, device=None, value_dtype="float64", shard=0):
    from liger_b200 import synthetic as syn
    Es, Ps, info = syn.make_torch(cfg["d"], n_per, cfg["m"], cfg["k"], density=cfg["density"], device=device,
                                  value_dtype=value_dtype, shard=shard)
    return Es, info


def workload_config(cfg, args, world):
    return {"workload": f"{args.config}: synthetic planted-Poisson iNMF, {cfg['d']} datasets x {cfg['n_per']} cells x "
                        f"{cfg['m']} genes per GPU, {int(cfg['density'] * 100)}% density, k={cfg['k']}, lambda={LAMBDA}",
            "datasets": cfg["d"], "cells_per_dataset_per_gpu": cfg["n_per"], "genes": cfg["m"], "k": cfg["k"],
            "density": cfg["density"], "total_cells": cfg["d"] * cfg["n_per"] * world, "parallelism": f"cell-shard x{world}",
            "l2_policy": "inputs larger than L2 (X is ~1.6 GB per GPU; no flush needed)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="C3")
    ap.add_argument("--cpu-fraction", type=float, default=0.5, help="cells fraction for the CPU legs")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl != "reference":
        args.warmup = max(args.warmup, 1)
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

    # ---- inputs: every rank generates ITS cell shard (weak scaling: n_per cells per dataset per GPU) ----
    t_gen = time.time()
    Es, info = make_inputs(cfg, cfg["n_per"], value_dtype="float64", shard=rank)
    torch.cuda.synchronize()
    t_gen = time.time() - t_gen
    d, m, k, n_per = cfg["d"], cfg["m"], cfg["k"], cfg["n_per"]
    N_local = d * n_per
    N_total = N_local * world
    nnz_local = info["nnz"]
    rng = liger_b200.RRandom(1)
    W0 = np.abs(rng.runif(m * k, 0, 2)).reshape((m, k), order="F")
    V0 = [np.abs(rng.runif(m * k, 0, 2)).reshape((m, k), order="F") for _ in range(d)]

    # ---- resident context.  Multi-GPU: one process per GPU, each passing ITS cell shard (LGR_FLAG_LOCAL_SHARD);
    #      the only collective is the library's per-iteration NCCL all-reduce of the W/V normal equations. ----
    nccl_id = None
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(liger_b200.nccl_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().numpy().tobytes())
    ctx = liger_b200.Context(Es, k, LAMBDA, W0, V0, device=local_rank, rank=rank, world=world, nccl_id=nccl_id,
                             flags=0x8 if world > 1 else 0)
    ctx.set_kernel_timing(True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up + timed region ----
    ctx.iterate(args.warmup)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    barrier()
    t0 = time.time()
    ctx.iterate(args.steps)
    barrier()
    t1 = time.time()
    tm = ctx.timings()
    kt = ctx.kernel_times()
    clocks = sampler.stop(t0, t1)
    loop_ms = torch.tensor([tm["loop_ms"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(loop_ms, op=dist.ReduceOp.MAX)
    loop_ms = float(loop_ms.item())
    obj = ctx.objective()
    value = N_total * args.steps / (loop_ms * 1e-3)

    # ---- roofline of the dominant kernel (CUDA events recorded on the library's launch stream) ----
    ab = algorithmic_bytes(cfg, nnz_local, N_local)
    dom = max(kt.items(), key=lambda kv: kv[1]["ms"])[0] if kt else None
    roofline = None
    if dom is not None:
        per_launch_ms = kt[dom]["ms"] / max(kt[dom]["launches"], 1)
        bytes_launch = ab.get(dom)
        ach = (bytes_launch / (per_launch_ms * 1e-3) / 1e9) if bytes_launch else None
        roofline = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": (ach / hbm_peak) if ach else None,
                    "traffic": ncu_traffic(dom) if args.config == "C3" else None, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": bytes_launch, "ms_per_launch": per_launch_ms,
                    "iteration": {"algorithmic_bytes": ab["iteration"],
                                  "achieved": ab["iteration"] / (loop_ms * 1e-3 / args.steps) / 1e9,
                                  "frac": ab["iteration"] / (loop_ms * 1e-3 / args.steps) / 1e9 / hbm_peak},
                    "kernels_ms_per_iter": {kn: v["ms"] / args.steps for kn, v in kt.items()}}
    ctx.close()

    # ---- e2e: lgr_inmf with HOST buffers (upload + K iterations from a cold start + download), through the public
    #      API at N GPUs: every rank passes its own pinned host shard; timed barrier-to-barrier, max over ranks;
    #      median of 3 calls (the PCIe copy rate of a fresh box varies from call to call) ----
    e2e = None
    if not args.no_e2e:
        def call(niter):
            return liger_b200.inmf(Es, k=k, lambda_=LAMBDA, niter=niter, Winit=W0, Vinit=V0, pinned_outputs=True,
                                   device=local_rank, rank=rank, world=world, nccl_id=nccl_id, local_shard=True)   # same unique id => the library reuses its communicator
        _w = call(1)   # warm the memory pools
        del _w
        runs = []
        for _rep in range(3):
            barrier()
            t0e = time.time()
            res = call(args.steps)
            barrier()
            dt = torch.tensor([time.time() - t0e], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            runs.append((float(dt.item()), res["timings"], res["objErr"]))
            h2d = sum(e[0].nbytes + e[1].nbytes + e[2].nbytes for e in Es) + W0.nbytes * (d + 1)
            d2h = sum(h.nbytes for h in res["H"]) + W0.nbytes * (d + 1)
            del res
        runs.sort(key=lambda r: r[0])
        dt, tms, obj_e = runs[1]
        e2e = {"value": N_total * args.steps / dt, "unit": "cell*iter/s", "h2d_bytes_per_step": h2d * world / args.steps,
               "d2h_bytes_per_step": d2h * world / args.steps, "seconds": dt, "seconds_all_calls": [r[0] for r in runs],
               "timings_ms": tms, "timings_ms_all_calls": [r[1] for r in runs], "objErr": obj_e,
               "note": "one lgr_inmf call per GPU = H2D of that rank's X shard (dgCMatrix slots, f64 values, pinned) + inits, "
                       "device layout build, K ANLS iterations from a cold start, objective, D2H of W/V/H (f64, pinned); "
                       "barrier-to-barrier, max over ranks, median of 3 calls"}

    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import c_oracle
        # bounded sample: every cell of the workload when that is 10-30 s of CPU work (C3 on 16 threads: ~13 s), else a
        # prefix of each dataset
        frac = 1.0 if (args.config == "C3" and args.cpu_fraction == 0.5) else args.cpu_fraction
        n_s = max(int(n_per * frac), k + 1)
        mats = [syn.to_scipy((e[0][:n_s + 1], e[1][:e[0][n_s]], e[2][:e[0][n_s]], (m, n_s))) for e in Es]
        nth = c_oracle.num_threads()
        it_cpu = 30
        tc0 = time.time()
        c_oracle.inmf(mats, k, LAMBDA, it_cpu, W0, V0, nthreads=nth)
        dtc = time.time() - tc0
        cpu = {"value": d * n_s * it_cpu / dtc, "unit": "cell*iter/s", "cores": nth, "kind": "port",
               "sample": f"first {n_s} cells of each of the {d} datasets, {it_cpu} iterations from the same init, FP64, "
                         f"OpenMP x{nth} ({dtc:.1f} s); CPU restatement of RcppPlanc inmf (oracle/inmf_oracle.c)"}

    if rank == 0:
        line = {"metric": "iNMF cell*iterations/s", "value": value, "unit": "cell*iter/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": loop_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32 (fp64 Gram / V,W solves / objective)",
                "data": "synthetic", "config": workload_config(cfg, args, world), "clocks": clocks,
                "e2e": e2e, "gpu_launches": int(tm["kernel_launches"]), "roofline": roofline, "cpu_baseline": cpu,
                "objErr": obj, "nnz_per_gpu": nnz_local, "gen_seconds": t_gen,
                "host_cpus_bound": (f"{len(cpus)} cores local to the GPU" if cpus else None),
                "nnls": {"pivots": tm["nnls_pivots"], "unconverged": tm["nnls_unconverged"]}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
