"""liger_b200 -- B200-native iNMF / UINMF alternating-NNLS path of rliger (drop-in for the
RcppPlanc::inmf / uinmf / bppnnls_prod calls and src/objErr.cpp).  See DESIGN.md."""
from .api import online_inmf, inmf, uinmf, bppnnls_prod, objErr_i, Context, RRandom, nccl_unique_id, device_count, CountScaled  # noqa: F401
from .integration import run_inmf_list, run_uinmf_list, runINMF, runUINMF  # noqa: F401

__version__ = "0.1.0"
