"""Developer check: where the wall time of one lgr_inmf call goes (C3, pinned inputs/outputs)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, liger_b200, torch
from liger_b200 import api, _ffi
import ctypes as C
cfg = bench.CONFIGS["C3"]
Es, info = bench.make_inputs(cfg, cfg["n_per"])
m, k, d = cfg["m"], cfg["k"], cfg["d"]
rng = liger_b200.RRandom(1)
W0 = np.abs(rng.runif(m * k, 0, 2)).reshape((m, k), order="F")
V0 = [np.abs(rng.runif(m * k, 0, 2)).reshape((m, k), order="F") for _ in range(d)]
mats = [liger_b200.CountScaled(e, rs, cs) for e, (rs, cs) in zip(Es, info["scales"])]   # the dense-tile path, as bench.py runs it
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    a = api._Args(mats, k, 5.0, 30, W0, V0, None, flags=0, device=-1)
    t1 = time.time()
    o = api._Out(a, pinned=True)
    t2 = time.time()
    _ffi.check(_ffi.lib().lgr_inmf(C.byref(a.struct), C.byref(o.struct)))
    t3 = time.time()
    tm = o.timings()
    del o, a
    t4 = time.time()
    print(f"args {1e3*(t1-t0):.1f} ms | out alloc {1e3*(t2-t1):.1f} | lgr_inmf {1e3*(t3-t2):.1f} (upload {tm['upload_ms']:.1f} loop {tm['loop_ms']:.1f} download {tm['download_ms']:.1f}) | free {1e3*(t4-t3):.1f}")
