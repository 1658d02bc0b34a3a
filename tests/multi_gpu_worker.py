"""Worker for tests/test_gpu_multi.py: one rank per GPU, cell-sharded iNMF through the C ABI (world = WORLD_SIZE)."""
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import liger_b200  # noqa: E402
from oracle import inmf_oracle as orc  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = np.load(os.path.join(ROOT, "tests", "golden", "pbmc_golden.npz"))
    Es = [sp.csc_matrix((g[f"scale_{n}_data"], g[f"scale_{n}_indices"], g[f"scale_{n}_indptr"]),
                        shape=tuple(g[f"scale_{n}_shape"])) for n in ("ctrl", "stim")]
    W0, V0, _, _ = orc.seeded_init(1, 173, 20, [300, 300])
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.frombuffer(bytearray(liger_b200.nccl_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(idt, 0)
    nccl_id = bytes(idt.cpu().numpy().tobytes())
    mode = sys.argv[1]
    if mode.startswith("dense"):   # count-structured synthetic data: the dense-tile tensor-core sweeps on cell shards
        from liger_b200 import synthetic as syn
        d, n, m, k = 3, 2100, 700, int(mode[5:] or 24)   # "dense44": k > 32, the sweeps run once per half of the factors
        Es, _, sc = syn.make_numpy(d, n, m, 8, block=700, seed=21, return_scales=True)
        W0, V0, _, _ = orc.seeded_init(1, m, k, [n] * d)
        Cs = [liger_b200.CountScaled(E, rs, cs) for E, (rs, cs) in zip(Es, sc)]
        out = liger_b200.inmf(Cs, k=k, lambda_=5.0, niter=6, Winit=W0, Vinit=V0, device=local, rank=rank, world=world,
                              nccl_id=nccl_id, trajectory=True)     # slice mode: the library takes this rank's columns
        np.savez(os.path.join(sys.argv[2], f"{mode}_rank{rank}.npz"), W=out["W"], V0=out["V"][0], H0=out["H"][0], traj=out["traj"],
                 obj=out["objErr"])
        dist.destroy_process_group()
        return
    if mode == "datasets":   # whole-dataset placement (LGR_FLAG_DATASET_SHARD): rank r holds dataset r with all its cells
        assert world == 2
        ctx = liger_b200.Context([Es[rank]], 20, 5.0, W0, [V0[rank]], device=local, rank=rank, world=world, nccl_id=nccl_id,
                                 flags=0x80)
        traj = ctx.iterate(10, trajectory=True)
        res = ctx.download()
        obj = ctx.objective()
        ctx.close()
        # also the one-shot entry point, dense-tile sweeps on count-structured data, two datasets per rank
        from liger_b200 import synthetic as syn
        d, n, m, k = 4, 1500, 600, 16
        Xs, _, sc = syn.make_numpy(d, n, m, 8, block=500, seed=33, return_scales=True)
        Wd, Vd, _, _ = orc.seeded_init(1, m, k, [n] * d)
        mine = [2 * rank, 2 * rank + 1]
        Cs = [liger_b200.CountScaled(Xs[i], *sc[i]) for i in mine]
        out = liger_b200.inmf(Cs, k=k, lambda_=5.0, niter=5, Winit=Wd, Vinit=[Vd[i] for i in mine], device=local, rank=rank,
                              world=world, nccl_id=nccl_id, dataset_shard=True)
        np.savez(os.path.join(sys.argv[2], f"{mode}_rank{rank}.npz"), W=res["W"], V=res["V"][0], H=res["H"][0], traj=traj, obj=obj,
                 dW=out["W"], dV=out["V"][0], dH=out["H"][1], dobj=out["objErr"])
        dist.destroy_process_group()
        return
    if mode == "slice":      # every rank passes the FULL matrices; the library takes its column range
        ctx = liger_b200.Context(Es, 20, 5.0, W0, V0, device=local, rank=rank, world=world, nccl_id=nccl_id)
        lo = [0, 0]
    else:                    # every rank passes only its own cells (LGR_FLAG_LOCAL_SHARD)
        rng_ = [(E.shape[1] * rank // world, E.shape[1] * (rank + 1) // world) for E in Es]
        Xs = [E[:, a:b].tocsc() for E, (a, b) in zip(Es, rng_)]
        ctx = liger_b200.Context(Xs, 20, 5.0, W0, V0, device=local, rank=rank, world=world, nccl_id=nccl_id, flags=0x8)
        lo = [a for a, _ in rng_]
    traj = ctx.iterate(10, trajectory=True)
    res = ctx.download()
    obj = ctx.objective()
    ctx.close()
    np.savez(os.path.join(sys.argv[2], f"{mode}_rank{rank}.npz"), W=res["W"], V0=res["V"][0], V1=res["V"][1],
             H0=res["H"][0], H1=res["H"][1], traj=traj, obj=obj, lo=np.array(lo))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
