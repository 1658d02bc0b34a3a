"""CPU suite, world_size = 2 over gloo: the cell-sharded formulation of the path.

Every rank solves H for its own cells and contributes the k x k / m x k sufficient statistics
G_i = H_i^T H_i and R_i = X_i H_i; ONE sum all-reduce per iteration makes the V_i / W normal equations
global, and every rank then solves V_i and W redundantly.  This test runs exactly that recipe with the
FP64 oracle pieces on two gloo ranks and checks it reproduces the single-process oracle (and therefore the
reference's golden factors): it pins the host-side sharding logic the CUDA library implements
(column ranges [n*r/w, n*(r+1)/w), statistics packing, redundant solves)."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, rel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, niter, outdir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import inmf_oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = np.load(os.path.join(GOLDEN, "pbmc_golden.npz"))
    Es = [sp.csc_matrix((g[f"scale_{n}_data"], g[f"scale_{n}_indices"], g[f"scale_{n}_indptr"]),
                        shape=tuple(g[f"scale_{n}_shape"])) for n in ("ctrl", "stim")]
    m, k, lam = 173, 20, 5.0
    W, Vs, _, _ = orc.seeded_init(1, m, k, [E.shape[1] for E in Es])
    # this rank's cell shard of every dataset (same split rule as lgr_ctx_create)
    rng_ = [(E.shape[1] * rank // world, E.shape[1] * (rank + 1) // world) for E in Es]
    Xs = [E[:, a:b] for E, (a, b) in zip(Es, rng_)]
    Hs = [None, None]
    for _ in range(niter):
        stats = []
        for i, X in enumerate(Xs):
            Hs[i] = orc.solve_H(W, Vs[i], X, lam)                  # local cells only
            stats.append(Hs[i].T @ Hs[i])                           # G_i  (k x k)
            stats.append(np.asarray((X @ Hs[i])))                   # R_i  (m x k)
        flat = torch.from_numpy(np.concatenate([s.ravel() for s in stats]))
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)                 # the ONLY collective of the iteration
        flat = flat.numpy()
        off = 0
        G, R = [], []
        for i in range(len(Xs)):
            G.append(flat[off:off + k * k].reshape(k, k)); off += k * k
            R.append(flat[off:off + m * k].reshape(m, k)); off += m * k
        for i in range(len(Xs)):                                    # redundant V_i solves (old W)
            Vs[i] = orc.bppnnls_prod((1 + lam) * G[i], R[i].T - G[i] @ W.T).T
        CtC = sum(G)
        CtB = sum(R[i].T - G[i] @ Vs[i].T for i in range(len(Xs)))
        W = orc.bppnnls_prod(CtC, CtB).T                            # redundant W solve (new V_i)
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), W=W, V0=Vs[0], V1=Vs[1], H0=Hs[0], H1=Hs[1],
             a0=rng_[0][0], a1=rng_[1][0])
    dist.destroy_process_group()


def test_cell_sharded_anls_world2_gloo(tmp_path, pbmc):
    from oracle import inmf_oracle as orc
    niter = 6
    port = _free_port()
    mp.spawn(_worker, args=(2, port, niter, str(tmp_path)), nprocs=2, join=True)
    W0, V0, _, _ = orc.seeded_init(1, 173, 20, [300, 300])
    ref = orc.inmf(pbmc["Es"], 20, 5.0, niter, W0, V0)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    # replicated factors are identical on both ranks and equal the single-process result
    assert np.array_equal(r0["W"], r1["W"]) and np.array_equal(r0["V0"], r1["V0"])
    assert rel(r0["W"], ref["W"]) < 1e-10 and rel(r0["V1"], ref["V"][1]) < 1e-10
    # H shards concatenate to the full H
    for i, key in enumerate(("H0", "H1")):
        H = np.vstack([r0[key], r1[key]])
        assert H.shape == ref["H"][i].shape and rel(H, ref["H"][i]) < 1e-10


def _worker_datasets(rank, world, port, niter, outdir):
    """Whole-dataset placement (LGR_FLAG_DATASET_SHARD): rank r owns dataset r with all its cells; H_r and V_r are solved
    locally and the ONLY collective of an iteration is the sum of the W solve's normal equations."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import inmf_oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = np.load(os.path.join(GOLDEN, "pbmc_golden.npz"))
    Es = [sp.csc_matrix((g[f"scale_{n}_data"], g[f"scale_{n}_indices"], g[f"scale_{n}_indptr"]),
                        shape=tuple(g[f"scale_{n}_shape"])) for n in ("ctrl", "stim")]
    m, k, lam = 173, 20, 5.0
    W, Vs, _, _ = orc.seeded_init(1, m, k, [E.shape[1] for E in Es])
    X, V = Es[rank], Vs[rank]                                      # this rank's dataset, whole
    obj = 0.0
    for _ in range(niter):
        H = orc.solve_H(W, V, X, lam)
        G = H.T @ H
        R = np.asarray(X @ H)
        V = orc.bppnnls_prod((1 + lam) * G, R.T - G @ W.T).T        # local V solve (old W): no communication
        eq = torch.from_numpy(np.concatenate([G.ravel(), (R.T - G @ V.T).ravel()]))
        dist.all_reduce(eq, op=dist.ReduceOp.SUM)                   # sum_i G_i and sum_i (R_i - G_i V_i^T)^T: (k + m) x k doubles
        eq = eq.numpy()
        W = orc.bppnnls_prod(eq[:k * k].reshape(k, k), eq[k * k:].reshape(k, m)).T
    o = torch.tensor([orc.objerr_i(orc.solve_H(W, V, X, lam).T, W, V, X, lam)], dtype=torch.float64)
    dist.all_reduce(o, op=dist.ReduceOp.SUM)                        # objective: one scalar at the end
    np.savez(os.path.join(outdir, f"ds_rank{rank}.npz"), W=W, V=V, H=H, obj=float(o[0]))
    dist.destroy_process_group()


def test_dataset_placement_anls_world2_gloo(tmp_path, pbmc):
    from oracle import inmf_oracle as orc
    niter = 6
    port = _free_port()
    mp.spawn(_worker_datasets, args=(2, port, niter, str(tmp_path)), nprocs=2, join=True)
    W0, V0, _, _ = orc.seeded_init(1, 173, 20, [300, 300])
    ref = orc.inmf(pbmc["Es"], 20, 5.0, niter, W0, V0)
    r = [np.load(tmp_path / f"ds_rank{i}.npz") for i in range(2)]
    assert np.array_equal(r[0]["W"], r[1]["W"]) and r[0]["obj"] == r[1]["obj"]
    assert rel(r[0]["W"], ref["W"]) < 1e-10
    for i in range(2):
        assert rel(r[i]["V"], ref["V"][i]) < 1e-10 and rel(r[i]["H"], ref["H"][i]) < 1e-10
