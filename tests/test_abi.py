"""CPU suite: the C ABI library loads, exports every symbol the header declares, and fails loudly without a GPU."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import ROOT, have_gpu


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "liger_b200.h")).read()
    return sorted(set(re.findall(r"^LGR_API [^;(]*?\b(lgr_[a-z0-9_]+)\(", hdr, flags=re.M)))


def test_library_exports_every_declared_symbol():
    from liger_b200 import _ffi
    syms = _declared_symbols()
    assert len(syms) >= 22
    assert sorted(_ffi.EXPORTS) == syms
    lib = ctypes.CDLL(_ffi.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), s
    assert _ffi.lib().lgr_version() == 1


def test_struct_layouts_match_header():
    from liger_b200 import _ffi
    # sizes computed by hand from include/liger_b200.h (LP64)
    assert ctypes.sizeof(_ffi.lgr_csc) == 64
    assert ctypes.sizeof(_ffi.lgr_timings) == 72
    assert ctypes.sizeof(_ffi.lgr_inmf_args) == 120
    assert ctypes.sizeof(_ffi.lgr_inmf_out) == 56 + 72


@pytest.mark.skipif(have_gpu(), reason="checks the no-device error path")
def test_no_cpu_fallback():
    import liger_b200
    from liger_b200._ffi import LigerB200Error
    E = sp.random(30, 40, 0.3, format="csc", random_state=0)
    with pytest.raises(LigerB200Error) as e:
        liger_b200.inmf([E], k=5, lambda_=5.0, niter=1, Winit=np.ones((30, 5)), Vinit=[np.ones((30, 5))])
    assert e.value.code == 2 and "no CPU path" in str(e.value)
    with pytest.raises(LigerB200Error):
        liger_b200.bppnnls_prod(np.eye(3), np.ones((3, 2)))
    with pytest.raises(LigerB200Error):
        liger_b200.objErr_i(np.ones((5, 40)), np.ones((30, 5)), np.ones((30, 5)), E, 5.0)
    from liger_b200 import preprocess as pp
    for call in (lambda: pp.normalize(E), lambda: pp.rowMeansVars(E),
                 lambda: pp.scaleNotCenter(E, [str(i) for i in range(30)], ["1", "2"])):
        with pytest.raises(LigerB200Error) as e:
            call()
        assert e.value.code == 2


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "liger_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("inmf_oracle.c)", ""), f"{f} mentions the oracle"


def test_wrapper_validation_messages():
    """Same checks and messages as .runINMF.list (R/integration.R:398-403) / runINMF.liger (:247-251)."""
    import liger_b200
    E = sp.random(30, 12, 0.3, format="csc", random_state=0)
    with pytest.raises(ValueError, match="number of cells in the smallest dataset"):
        liger_b200.run_inmf_list([E, E], k=20)
    E2 = sp.random(8, 40, 0.3, format="csc", random_state=0)
    with pytest.raises(ValueError, match="number of shared features"):
        liger_b200.run_inmf_list([E2, E2], k=10)
    with pytest.raises(ValueError, match="Scaled data not available"):
        liger_b200.runINMF({"a": E, "b": None}, k=5)
    with pytest.raises(ValueError, match="No scaled data for unshared feature"):
        liger_b200.run_uinmf_list([E, E], [None, None], k=5)


def _build_cabi(tmp_path):
    """gcc (plain C, -std=c99 -pedantic) against include/liger_b200.h, linked to the in-tree shared library."""
    from liger_b200 import _ffi
    exe = str(tmp_path / "test_cabi")
    libdir = os.path.dirname(_ffi.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "test_cabi.c"), "-o", exe, "-L", libdir, "-lliger_b200",
                    f"-Wl,-rpath,{libdir}", "-lm"], check=True)
    return exe


def test_c_caller_without_gpu(tmp_path):
    """The ABI from C (not ctypes): header compiles as C99, argument errors, R RNG known answer, no-device error."""
    exe = _build_cabi(tmp_path)
    r = subprocess.run([exe, "nogpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cabi nogpu ok" in r.stdout


@pytest.mark.gpu
def test_c_caller_on_gpu(tmp_path):
    """A C program runs a small iNMF (explicit and in-library inits), bppnnls_prod and objErr_i through the ABI and
    sees malformed row indices rejected with LGR_ERR_INVALID."""
    exe = _build_cabi(tmp_path)
    r = subprocess.run([exe, "gpu"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cabi gpu ok" in r.stdout


def test_uinmf_list_pairing_is_validated():
    import liger_b200
    E = sp.random(30, 40, 0.3, format="csc", random_state=0)
    with pytest.raises(ValueError, match="unsharedList is a dict"):
        liger_b200.run_uinmf_list([E, E], {"a": E}, k=5)
    with pytest.raises(ValueError, match="one entry"):
        liger_b200.run_uinmf_list([E, E], [E], k=5)
