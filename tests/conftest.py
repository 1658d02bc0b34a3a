import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The CUDA library is built in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    from liger_b200 import build as b
    b.build()
    return b.LIB


def _csc(npz, name):
    return sp.csc_matrix((npz[f"scale_{name}_data"], npz[f"scale_{name}_indices"], npz[f"scale_{name}_indptr"]),
                         shape=tuple(npz[f"scale_{name}_shape"]))


@pytest.fixture(scope="session")
def pbmc():
    g = np.load(os.path.join(GOLDEN, "pbmc_golden.npz"))
    r = np.load(os.path.join(GOLDEN, "pbmc_raw.npz"))
    raws = [sp.csc_matrix((r[f"raw_{n}_data"], r[f"raw_{n}_indices"], r[f"raw_{n}_indptr"]), shape=tuple(r[f"raw_{n}_shape"]))
            for n in ("ctrl", "stim")]
    return {"Es": [_csc(g, "ctrl"), _csc(g, "stim")], "golden": g, "names": ("ctrl", "stim"), "raws": raws,
            "genes": [list(r["genes_ctrl"]), list(r["genes_stim"])], "nUMI": [r["nUMI_ctrl"], r["nUMI_stim"]]}


@pytest.fixture(scope="session")
def bmmc():
    g = np.load(os.path.join(GOLDEN, "bmmc_inputs.npz"))
    return {"Es": [_csc(g, "rna"), _csc(g, "atac")], "names": ("rna", "atac")}


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))


def pattern_mismatches(x, ref, margin=1e-5):
    """Zero-pattern disagreements that are not degenerate: one side is exactly 0 while the other is clearly
    non-zero (> margin * max|ref|).  (SURVEY 8d parity gate.)"""
    x = np.asarray(x)
    ref = np.asarray(ref)
    scale = margin * np.abs(ref).max()
    bad = ((x == 0) & (ref > scale)) | ((ref == 0) & (x > scale))
    return int(bad.sum())


def have_gpu():
    try:
        import liger_b200
        return liger_b200.device_count() > 0
    except Exception:
        return False
