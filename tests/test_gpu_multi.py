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
