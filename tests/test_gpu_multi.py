"""GPU suite, 2 GPUs: the cell-sharded path (one process per GPU, one NCCL all-reduce per iteration) must reproduce
the single-GPU / oracle result.  Skipped on a single-GPU box."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, rel
from oracle import inmf_oracle as orc

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("mode", ["slice", "local"])
def test_two_gpu_cell_sharding(tmp_path, pbmc, mode):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_worker.py"), mode, str(tmp_path)]
    subprocess.run(cmd, check=True, timeout=600, cwd=ROOT)
    W0, V0, _, _ = orc.seeded_init(1, 173, 20, [300, 300])
    ref = orc.inmf(pbmc["Es"], 20, 5.0, 10, W0, V0, trajectory=True)
    r0 = np.load(tmp_path / f"{mode}_rank0.npz")
    r1 = np.load(tmp_path / f"{mode}_rank1.npz")
    assert np.array_equal(r0["W"], r1["W"]) and np.array_equal(r0["V1"], r1["V1"])      # replicated factors identical
    assert rel(r0["W"], ref["W"]) < 1e-3 and rel(r0["V0"], ref["V"][0]) < 1e-3
    assert np.max(np.abs(r0["traj"] - np.array(ref["traj"])) / np.array(ref["traj"])) < 1e-4
    assert abs(float(r0["obj"]) - ref["objErr"]) < 1e-4 * ref["objErr"]
    for i, key in enumerate(("H0", "H1")):
        if mode == "slice":     # each rank wrote its rows into a full-size array
            H = r0[key] + r1[key]
        else:
            H = np.vstack([r0[key], r1[key]])
        assert rel(H, ref["H"][i]) < 1e-3


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("k", [24, 44])
def test_two_gpu_dense_sweeps(tmp_path, k):
    """The dense-tile tensor-core sweeps on cell shards (slice mode, one all-reduce per iteration) against the C oracle;
    k = 44: two passes per sweep (factor halves) and the k > 32 H-solve kernels."""
    from liger_b200 import synthetic as syn
    from oracle import c_oracle
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29534", os.path.join(ROOT, "tests", "multi_gpu_worker.py"), f"dense{k}", str(tmp_path)]
    subprocess.run(cmd, check=True, timeout=600, cwd=ROOT)
    d, n, m = 3, 2100, 700
    Es, _ = syn.make_numpy(d, n, m, 8, block=700, seed=21)
    W0, V0, _, _ = orc.seeded_init(1, m, k, [n] * d)
    ref = c_oracle.inmf(Es, k, 5.0, 6, W0, V0, trajectory=True)
    r0 = np.load(tmp_path / f"dense{k}_rank0.npz")
    r1 = np.load(tmp_path / f"dense{k}_rank1.npz")
    assert np.array_equal(r0["W"], r1["W"]) and np.array_equal(r0["V0"], r1["V0"])
    assert rel(r0["W"], ref["W"]) < 1e-3 and rel(r0["V0"], ref["V"][0]) < 1e-3
    assert rel(r0["H0"] + r1["H0"], ref["H"][0]) < 1e-3       # slice mode: each rank wrote its own rows
    assert np.max(np.abs(r0["traj"] - ref["traj"]) / ref["traj"]) < 1e-4


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_two_gpu_dataset_placement(tmp_path, pbmc):
    """LGR_FLAG_DATASET_SHARD: rank r holds dataset r entirely; the only collective is the all-reduce of the W solve's
    normal equations.  W / objective identical on both ranks and equal to the oracle's; V_r, H_r local."""
    from liger_b200 import synthetic as syn
    from oracle import c_oracle
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29535", os.path.join(ROOT, "tests", "multi_gpu_worker.py"), "datasets", str(tmp_path)]
    subprocess.run(cmd, check=True, timeout=600, cwd=ROOT)
    W0, V0, _, _ = orc.seeded_init(1, 173, 20, [300, 300])
    ref = orc.inmf(pbmc["Es"], 20, 5.0, 10, W0, V0, trajectory=True)
    r = [np.load(tmp_path / f"datasets_rank{i}.npz") for i in range(2)]
    assert np.array_equal(r[0]["W"], r[1]["W"]) and float(r[0]["obj"]) == float(r[1]["obj"])
    assert np.array_equal(r[0]["traj"], r[1]["traj"])
    assert rel(r[0]["W"], ref["W"]) < 1e-3
    assert np.max(np.abs(r[0]["traj"] - np.array(ref["traj"])) / np.array(ref["traj"])) < 1e-4
    assert abs(float(r[0]["obj"]) - ref["objErr"]) < 1e-4 * ref["objErr"]
    for i in range(2):
        assert rel(r[i]["V"], ref["V"][i]) < 1e-3 and rel(r[i]["H"], ref["H"][i]) < 1e-3
    d, n, m, k = 4, 1500, 600, 16
    Xs, _ = syn.make_numpy(d, n, m, 8, block=500, seed=33)
    Wd, Vd, _, _ = orc.seeded_init(1, m, k, [n] * d)
    refd = c_oracle.inmf(Xs, k, 5.0, 5, Wd, Vd)
    assert np.array_equal(r[0]["dW"], r[1]["dW"]) and rel(r[0]["dW"], refd["W"]) < 1e-3
    assert abs(float(r[0]["dobj"]) - refd["objErr"]) < 1e-4 * refd["objErr"] and float(r[0]["dobj"]) == float(r[1]["dobj"])
    for i in range(2):
        assert rel(r[i]["dV"], refd["V"][2 * i]) < 1e-3 and rel(r[i]["dH"], refd["H"][2 * i + 1]) < 1e-3


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_one_process_two_devices(pbmc):
    """The opt-in shared-memory attributes of the kernels are per device: a process that uses device 0 and then device 1
    must get the same factors from both (k = 20 and k = 40 cover the register, warp and shared-memory NNLS kernels)."""
    import liger_b200
    for k in (20, 40):
        W0, V0, _, _ = orc.seeded_init(1, 173, k, [300, 300])
        a = liger_b200.inmf(pbmc["Es"], k=k, lambda_=5.0, niter=4, Winit=W0, Vinit=V0, device=0)
        b = liger_b200.inmf(pbmc["Es"], k=k, lambda_=5.0, niter=4, Winit=W0, Vinit=V0, device=1)
        assert rel(a["W"], b["W"]) < 1e-5 and rel(a["H"][0], b["H"][0]) < 1e-5
    G = np.eye(40) + 0.1
    B = np.random.default_rng(0).random((40, 500))
    assert rel(liger_b200.bppnnls_prod(G, B, device=1), liger_b200.bppnnls_prod(G, B, device=0)) < 1e-12
