"""Mirror of the three R-level block solves of the reference (R/cINMF.R:477-503) and of the "consensus fixed"
refinement loop of runCINMF that is built from them (R/cINMF.R:392-407).  In R each solve is host algebra for the normal
equations followed by RcppPlanc::bppnnls_prod.  Here they are the library's own block solves on a RESIDENT context
(lgr_ctx_solve): the sparse matrices are uploaded once, the right-hand sides (W+V)^T E and E H are formed by the same
sweeps over X as inside lgr_inmf, the Grams and the NNLS run on the GPU, and only the factor a caller asks for comes
back.  The objective is objErr_i (src/objErr.cpp:13-31) from the same resident state (lgr_ctx_objerr_i).

    inmf_solveH(H, W, V, E, lambda)     -> n x k      R/cINMF.R:477-483
    inmf_solveV(H, W, V, E, lambda)     -> m x k      R/cINMF.R:485-491
    inmf_solveW(Hs, W, Vs, Es, lambda)  -> m x k      R/cINMF.R:493-503
    BlockSolver(Es, k, lambda, W, Vs, Hs=None) : the resident form -- solveH() / solveV() / solveW() / objErr() on data
        that stays in HBM between calls (what a loop over the three functions should use)
    refine_fixed_W(Es, W, Vs, lambda, nIteration, seed) : the loop of R/cINMF.R:392-407 (W fixed to the consensus,
        H_i / V_i alternate 2 * nIteration times), returning dict(H, V, W, objErr)
"""
import numpy as np

from . import api


class BlockSolver:
    """X_1..X_d resident in HBM; the three block solves and the objective act on the factors kept beside them."""

    def __init__(self, Es, k, lambda_, W, Vs, Hs=None, device=-1):
        self.ctx = api.Context(list(Es), k, lambda_, np.asarray(W, dtype=np.float64),
                               [np.asarray(v, dtype=np.float64) for v in Vs],
                               Hinit=None if Hs is None else [np.asarray(h, dtype=np.float64) for h in Hs], device=device)

    def set(self, W=None, Vs=None, Hs=None):
        self.ctx.set_factors(W=W, V=Vs, H=Hs)

    def solveH(self):                       # every dataset: R/cINMF.R:477-483
        self.ctx.solve("H")
        return self.ctx.download()["H"]     # list of n_i x k

    def solveV(self):                       # every dataset, with the current W: R/cINMF.R:485-491
        self.ctx.solve("V")
        return self.ctx.download()["V"]

    def solveW(self):                       # R/cINMF.R:493-503
        self.ctx.solve("W")
        return self.ctx.download()["W"]

    def objErr(self, i=None):
        return self.ctx.objective() if i is None else self.ctx.objErr_i(i)

    def factors(self):
        return self.ctx.download()

    def close(self):
        self.ctx.close()


def inmf_solveH(H, W, V, E, lambda_, nCores=2):
    s = BlockSolver([E], W.shape[1], lambda_, W, [V])
    try:
        return s.solveH()[0]                # cell x factor
    finally:
        s.close()


def inmf_solveV(H, W, V, E, lambda_, nCores=2):
    V0 = np.zeros_like(W) if V is None else V
    s = BlockSolver([E], W.shape[1], lambda_, W, [V0], Hs=[H])
    try:
        return s.solveV()[0]
    finally:
        s.close()


def inmf_solveW(Hs, W, Vs, Es, lambda_, nCores=2):
    W0 = np.zeros_like(Vs[0]) if W is None else W
    s = BlockSolver(Es, Vs[0].shape[1], lambda_, W0, Vs, Hs=Hs)
    try:
        return s.solveW()
    finally:
        s.close()


def refine_fixed_W(Es, W, Vs, lambda_=5.0, nIteration=30, seed=1, nCores=2):
    """ANLS with the consensus W fixed (R/cINMF.R:386-407).  The H inits are drawn (and immediately overwritten by the
    first H solve) exactly like the reference does, so the RNG stream stays aligned with R."""
    k = W.shape[1]
    rng = api.RRandom(seed)
    for E in Es:
        rng.runif(E.shape[1] * k, 0.0, 2.0)
    s = BlockSolver(Es, k, lambda_, W, Vs)
    try:
        for _ in range(2 * nIteration):
            s.ctx.solve("HV")               # all H_i, then all V_i with the fixed W
        f = s.factors()
        return {"H": f["H"], "V": f["V"], "W": np.asarray(W, dtype=np.float64), "objErr": s.objErr()}
    finally:
        s.close()
