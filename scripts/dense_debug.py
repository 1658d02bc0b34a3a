"""Developer check of the dense-tile sweeps: each sweep alone (LGR_DENSE_MASK) against the gather path, one iteration."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1:
    import numpy as np
    import liger_b200
    from liger_b200 import synthetic as syn
    from oracle import inmf_oracle as orc
    d, n, m, k = [int(x) for x in sys.argv[2:6]]
    Es, _, sc = syn.make_numpy(d, n, m, 8, block=1000, seed=9, return_scales=True)
    Cs = [liger_b200.CountScaled(E, a, b) for E, (a, b) in zip(Es, sc)]
    W0, V0, _, _ = orc.seeded_init(1, m, k, [n] * d)
    out = liger_b200.inmf(Cs, k=k, lambda_=5.0, niter=int(sys.argv[6]), Winit=W0, Vinit=V0)
    np.savez(sys.argv[1], W=out["W"], H=out["H"][0], V=out["V"][0], obj=out["objErr"])
    print("launches", out["timings"]["kernel_launches"], "obj", out["objErr"])
else:
    import numpy as np
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
    for shape in (["2", "300", "100", "8", "1"], ["2", "3000", "2000", "30", "2"]):
        res = {}
        for mask in ("0", "1", "2", "3"):
            env = dict(os.environ, LGR_DENSE_MASK=mask)
            f = f"/tmp/dd_{mask}.npz"
            r = subprocess.run(["timeout", "120", sys.executable, __file__, f] + shape, env=env, capture_output=True, text=True)
            print("mask", mask, "rc", r.returncode, r.stdout.strip()[-200:], r.stderr.strip()[-300:])
            if r.returncode == 0:
                res[mask] = np.load(f)
        for mask in ("1", "2", "3"):
            if mask in res and "0" in res:
                print(shape, "mask", mask, "H", rel(res[mask]["H"], res["0"]["H"]), "W", rel(res[mask]["W"], res["0"]["W"]),
                      "V", rel(res[mask]["V"], res["0"]["V"]), "obj", float(res[mask]["obj"]), float(res["0"]["obj"]))
