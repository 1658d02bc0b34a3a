"""CPU suite: the C ABI library loads, exports every symbol the header declares, and fails loudly without a GPU."""
import ctypes
import os
import re

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
    assert ctypes.sizeof(_ffi.lgr_csc) == 48
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
