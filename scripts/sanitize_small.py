"""Small end-to-end calls that touch every kernel family on tiny inputs: meant for `compute-sanitizer --tool memcheck|racecheck`
(closed on the round-1 GPU pool, so it was only run plainly there)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import liger_b200
from liger_b200 import synthetic as syn, preprocess as pp
from oracle import inmf_oracle as orc
for k, m, ns in ((20, 120, (300, 170)), (44, 130, (260, 190))):
    Es = [syn.make_numpy(1, n, m, 8, block=500, seed=3 + i)[0][0] for i, n in enumerate(ns)]
    W0, V0, _, _ = orc.seeded_init(2, m, k, list(ns))
    for kw in ({}, {"tensor_u": True}, {"legacy_nnls": True}):
        out = liger_b200.inmf(Es, k=k, lambda_=5.0, niter=3, Winit=W0, Vinit=V0, trajectory=True, **kw)
        assert np.isfinite(out["objErr"])
Ps = [Es[0][100:], None]
W0, V0, U0, _ = orc.seeded_init(3, 100, 30, [260, 190], [30, 0])
out = liger_b200.uinmf([E[:100] for E in Es], Ps, k=30, lambda_=[5.0, 2.0], niter=2, Winit=W0, Vinit=V0, Uinit=U0)
G = np.eye(12) + 0.2
liger_b200.bppnnls_prod(G, np.random.default_rng(0).random((12, 333)))
n = pp.normalize(Es[0]); pp.rowMeansVars(n); pp.scaleNotCenter(n, [str(i) for i in range(m)], [str(i) for i in range(5, 60)])
print("sanitize run ok")
