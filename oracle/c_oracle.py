"""TEST INFRASTRUCTURE ONLY (oracle/): builds and binds oracle/inmf_oracle.c (FP64 C + OpenMP).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it."""
import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "inmf_oracle.c")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "liboracle.so")
_lib = None


def build(force=False):
    os.makedirs(BUILD, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        # -march=x86-64-v3 (AVX2+FMA) rather than native: the .so is built here and runs on the GPU box's host
        cmd = ["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
        subprocess.run(cmd, check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        L.orc_bppnnls_prod.argtypes = [C.c_int, C.c_int64, dp, dp, dp, C.c_int]
        L.orc_inmf.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, ip, ip, C.POINTER(ip), C.POINTER(ip), C.POINTER(dp),
                               dp, dp, C.POINTER(dp), C.POINTER(dp), dp, dp, C.c_int, C.POINTER(C.c_int64)]
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def bppnnls_prod(CtC, CtB, nthreads=0):
    G = np.ascontiguousarray(CtC, dtype=np.float64)
    B = np.asfortranarray(CtB, dtype=np.float64)
    k, c = B.shape
    X = np.zeros((k, c), order="F")
    rc = lib().orc_bppnnls_prod(k, c, _dp(G), _dp(B), _dp(X), nthreads)
    if rc == 1:
        raise ValueError("bad k")
    return X


def inmf(Es, k, lams, niter, Winit, Vinit, Ps=None, Uinit=None, Hinit=None, nthreads=0, trajectory=False):
    """iNMF (Ps None) / UINMF on the C oracle.  Returns dict(H=[n_i x k], V, W, U, objErr, traj, pivots)."""
    d = len(Es)
    m = Es[0].shape[0]
    lam = np.broadcast_to(np.asarray(lams, dtype=np.float64), (d,)).copy()
    n = np.array([E.shape[1] for E in Es], dtype=np.int32)
    u = np.zeros(d, dtype=np.int32)
    X, Vst = [], []
    for i in range(d):
        E = sp.csc_matrix(Es[i], dtype=np.float64)
        V0 = np.asarray(Vinit[i], dtype=np.float64)
        if Ps is not None and Ps[i] is not None and Ps[i].shape[0] > 0:
            u[i] = Ps[i].shape[0]
            E = sp.vstack([E, sp.csc_matrix(Ps[i], dtype=np.float64)], format="csc")
            V0 = np.vstack([V0, np.asarray(Uinit[i], dtype=np.float64)])
        E.sort_indices()
        X.append(E)
        Vst.append(np.array(V0, dtype=np.float64, order="F", copy=True))
    W = np.array(Winit, dtype=np.float64, order="F", copy=True)
    H = [np.zeros((int(n[i]), k), order="F") if Hinit is None else np.array(Hinit[i], dtype=np.float64, order="F", copy=True)
         for i in range(d)]
    cp = [np.ascontiguousarray(x.indptr, dtype=np.int32) for x in X]
    ri = [np.ascontiguousarray(x.indices, dtype=np.int32) for x in X]
    vv = [np.ascontiguousarray(x.data, dtype=np.float64) for x in X]
    ipp = C.POINTER(C.c_int)
    dpp = C.POINTER(C.c_double)
    cpa = (ipp * d)(*[_ip(a) for a in cp])
    ria = (ipp * d)(*[_ip(a) for a in ri])
    vva = (dpp * d)(*[_dp(a) for a in vv])
    Va = (dpp * d)(*[_dp(a) for a in Vst])
    Ha = (dpp * d)(*[_dp(a) for a in H])
    traj = np.zeros(max(niter, 1))
    obj = C.c_double(0.0)
    piv = C.c_int64(0)
    rc = lib().orc_inmf(d, k, niter, m, _ip(n), _ip(u), cpa, ria, vva, _dp(lam), _dp(W), Va, Ha,
                        _dp(traj) if trajectory else None, C.byref(obj), nthreads, C.byref(piv))
    if rc == 1:
        raise ValueError("bad arguments")
    return {"H": H, "W": W, "V": [v[:m] for v in Vst], "U": [v[m:] if u[i] > 0 else None for i, v in enumerate(Vst)],
            "objErr": obj.value, "traj": traj[:niter] if trajectory else None, "pivots": piv.value, "status": rc}


def num_threads():
    return lib().orc_num_threads()
