"""GPU suite (-m gpu): the CUDA path, called through the C ABI, against the oracle / golden fixtures.

Tolerances are the north_star's: factors <= 1e-3 relative Frobenius, objective trajectory <= 1e-4 relative,
identical NNLS zero patterns away from degenerate coordinates.  (Measured errors are ~1e-6.)"""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import rel, pattern_mismatches
from oracle import inmf_oracle as orc
from oracle import c_oracle

pytestmark = pytest.mark.gpu

TOL_FACTOR = 1e-3
TOL_OBJ = 1e-4


def _check(out, ref, tol=TOL_FACTOR, u=False):
    assert rel(out["W"], ref["W"]) < tol
    for a, b in zip(out["V"], ref["V"]):
        assert rel(a, b) < tol
    for a, b in zip(out["H"], ref["H"]):
        assert a.shape == b.shape
        assert rel(a, b) < tol
        assert pattern_mismatches(a, b) == 0
        assert a.min() >= 0
    if u:
        for a, b in zip(out["U"], ref["U"]):
            assert (a is None) == (b is None)
            if a is not None:
                assert rel(a, b) < tol


# ----------------------------------------------------------------------------------------------
# bppnnls_prod  (RcppPlanc::bppnnls_prod; R/cINMF.R:481)
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k,c", [(1, 5), (5, 1), (20, 127), (20, 129), (30, 1000), (33, 300), (50, 4097), (64, 65),
                                 (8, 160000)])   # the last one is beyond the warp-per-column kernel's batch limit
def test_bppnnls_prod_matches_oracle(k, c):
    import liger_b200
    rng = np.random.default_rng(k * 1000 + c)
    C = rng.random((3 * k + 2, k))
    G = C.T @ C
    B = C.T @ (rng.random((3 * k + 2, c)) - 0.25)
    X = liger_b200.bppnnls_prod(G, B)
    Xo = orc.bppnnls_prod(G, B)
    assert X.shape == (k, c)
    assert rel(X, Xo) < 1e-10
    assert np.array_equal(X > 0, Xo > 0)


def test_bppnnls_prod_edge_cases():
    import liger_b200
    rng = np.random.default_rng(0)
    # all-negative rhs -> all zeros; all-positive diagonal -> unconstrained solution
    G = np.diag(rng.random(8) + 0.5)
    assert np.all(liger_b200.bppnnls_prod(G, -np.ones((8, 10))) == 0)
    B = rng.random((8, 10))
    assert np.allclose(liger_b200.bppnnls_prod(G, B), B / np.diag(G)[:, None], rtol=1e-12)
    # singular Gram (a dead factor: zero row/column) must not produce NaN; the live block is solved exactly
    C = rng.random((40, 6))
    C[:, 3] = 0.0
    G = C.T @ C
    B = C.T @ rng.random((40, 33))
    X = liger_b200.bppnnls_prod(G, B)
    assert np.all(np.isfinite(X)) and np.all(X[3] == 0)
    live = [0, 1, 2, 4, 5]
    Xo = orc.bppnnls_prod(G[np.ix_(live, live)], B[live])
    assert rel(X[live], Xo) < 1e-9
    # vector rhs, zero columns
    assert liger_b200.bppnnls_prod(np.eye(3), np.array([1.0, -1.0, 2.0])).ravel().tolist() == [1.0, 0.0, 2.0]
    assert liger_b200.bppnnls_prod(np.eye(3), np.zeros((3, 0))).shape == (3, 0)


# ----------------------------------------------------------------------------------------------
# objErr_i  (src/objErr.cpp:13-31)
# ----------------------------------------------------------------------------------------------
def test_objerr_i_matches_oracle(pbmc):
    import liger_b200
    rng = np.random.default_rng(5)
    for E in pbmc["Es"]:
        for k in (7, 20, 40):
            W, V, H = rng.random((173, k)), rng.random((173, k)), rng.random((k, 300))
            H[rng.random(H.shape) < 0.4] = 0
            got = liger_b200.objErr_i(H, W, V, E, 5.0)
            want = orc.objerr_i(H, W, V, E, 5.0)
            assert abs(got - want) < 2e-6 * abs(want)


# ----------------------------------------------------------------------------------------------
# inmf  (RcppPlanc::inmf; R/integration.R:438-441)
# ----------------------------------------------------------------------------------------------
def test_inmf_reproduces_reference_golden_factors(pbmc):
    """C1: runINMF(pbmc, k=20, lambda=5, nIteration=30, seed=1) vs data/pbmcPlot.rda."""
    import liger_b200
    g = pbmc["golden"]
    out = liger_b200.run_inmf_list(pbmc["Es"], k=20, lambda_=5.0, nIteration=30, seed=1, trajectory=True)
    rows = g["golden_rows"]
    assert rel(out["W"][rows], g["golden_W"]) < TOL_FACTOR
    for i, n in enumerate(pbmc["names"]):
        assert out["H"][i].shape == (20, 300)              # transposed like R/integration.R:453
        assert rel(out["V"][i][rows], g[f"golden_V_{n}"]) < TOL_FACTOR
        assert rel(out["H"][i], g[f"golden_H_{n}"]) < TOL_FACTOR
        assert pattern_mismatches(out["H"][i], g[f"golden_H_{n}"]) == 0
    ref = orc.run_inmf_list(pbmc["Es"], k=20, lam=5.0, niter=30, seed=1, trajectory=True)
    assert np.max(np.abs(out["traj"] - np.array(ref["traj"])) / np.array(ref["traj"])) < TOL_OBJ
    assert abs(out["objErr"] - 36105.069129349) < TOL_OBJ * 36105.07
    # measured headroom: keep an eye on drift (fp32 path is ~1e-6)
    assert rel(out["W"][rows], g["golden_W"]) < 5e-5


def test_inmf_reference_test_config(pbmc):
    """tests/testthat/test_factorization.R:49: k = 10, nIteration = 2 -> valid shapes; values vs oracle."""
    import liger_b200
    out = liger_b200.run_inmf_list(dict(zip(pbmc["names"], pbmc["Es"])), k=10, nIteration=2, trajectory=True)
    assert out["W"].shape == (173, 10)
    for n in pbmc["names"]:
        assert out["H"][n].shape == (10, 300) and out["V"][n].shape == (173, 10)
    assert np.allclose(out["traj"], [66551.145423, 49910.681818], rtol=TOL_OBJ)


def test_inmf_bmmc_trajectory(bmmc):
    """C2: bmmc (rna + atac), k = 20, objective trajectory."""
    import liger_b200
    W0, V0, _, _ = orc.seeded_init(1, 135, 20, [340, 340])
    out = liger_b200.inmf(bmmc["Es"], k=20, lambda_=5.0, niter=30, Winit=W0, Vinit=V0, trajectory=True)
    ref = orc.inmf(bmmc["Es"], 20, 5.0, 30, W0, V0, trajectory=True)
    _check(out, ref)
    assert np.max(np.abs(out["traj"] - np.array(ref["traj"])) / np.array(ref["traj"])) < TOL_OBJ
    assert np.all(np.diff(out["traj"]) < 0)


@pytest.mark.parametrize("k,ns,m", [(30, (333, 1000, 57), 300), (40, (200, 129), 150), (3, (40, 41, 42, 43), 64),
                                    (63, (300, 170), 160), (50, (2500, 900), 200)])
def test_inmf_synthetic_ragged(k, ns, m):
    """ragged datasets, k > 32 (two-line factor rows; k = 50 exercises the flagged-column hand-over between the
    register and the warp NNLS kernels, k = 63 is beyond both), many datasets; vs the C oracle from identical inits."""
    import liger_b200
    from liger_b200 import synthetic as syn
    Es = []
    for i, n in enumerate(ns):
        E, _ = syn.make_numpy(1, n, m, 8, block=500, seed=100 + i)
        Es.append(E[0])
    W0, V0, _, _ = orc.seeded_init(2, m, k, list(ns))
    out = liger_b200.inmf(Es, k=k, lambda_=2.5, niter=6, Winit=W0, Vinit=V0, trajectory=True)
    ref = c_oracle.inmf(Es, k, 2.5, 6, W0, V0, trajectory=True)
    _check(out, ref)
    assert np.max(np.abs(out["traj"] - ref["traj"]) / ref["traj"]) < TOL_OBJ


@pytest.mark.parametrize("name,d,n,m,k,lam,niter", [
    ("C3-mini", 8, 3000, 2000, 30, 5.0, 6),      # gt = 1000, ngt = 2, KP = 32: the geometry bench.py runs at C3
    ("C4-mini", 4, 2000, 3000, 50, 5.0, 6),      # KP = 64, gt = 500, ngt = 6, register -> warp NNLS hand-over (k = 50)
    ("C3-ragged", 3, 2777, 2000, 30, 2.0, 4),    # cell count not a multiple of any block size, other lambda
])
def test_inmf_multi_gene_tile_geometry(name, d, n, m, k, lam, niter):
    """The code path the benchmark runs: m > 1024 (k <= 32) / m > 512 (k > 32) makes k_spmm_ltx_tiled walk several gene
    tiles per cell block (mbarrier phase flips, partial B rows accumulated across tiles).  From identical inits against
    the FP64 C oracle, with the north-star gates: factors <= 1e-3, trajectory <= 1e-4, zero patterns."""
    import liger_b200
    from liger_b200 import synthetic as syn
    Es, _ = syn.make_numpy(d, n, m, k, block=1000, seed=31 + d)
    W0, V0, _, _ = orc.seeded_init(1, m, k, [n] * d)
    out = liger_b200.inmf(Es, k=k, lambda_=lam, niter=niter, Winit=W0, Vinit=V0, trajectory=True)
    ref = c_oracle.inmf(Es, k, lam, niter, W0, V0, trajectory=True)
    _check(out, ref)
    assert np.max(np.abs(out["traj"] - ref["traj"]) / ref["traj"]) < TOL_OBJ
    assert abs(out["objErr"] - ref["objErr"]) < TOL_OBJ * ref["objErr"]
    assert out["timings"]["nnls_unconverged"] == 0


def test_uinmf_c5_shape_in_miniature():
    """C5 in miniature: 2 datasets, 2000 shared + 1000 unshared genes each, k = 30, per-dataset lambda -- stacked rows
    (3000) walk three gene tiles; UINMF parity is unpinned in the reference, so: oracle from identical inits + the
    documented objective (R/integration.R:1109-1113) decreasing monotonically."""
    import liger_b200
    from liger_b200 import synthetic as syn
    d, n, m, u, k = 2, 2500, 2000, 1000, 30
    Es, Ps = syn.make_numpy(d, n, m, k, u=u, block=1250, seed=77)
    lam = [5.0, 2.0]
    W0, V0, U0, _ = orc.seeded_init(1, m, k, [n] * d, [u] * d)
    out = liger_b200.uinmf(Es, Ps, k=k, lambda_=lam, niter=6, Winit=W0, Vinit=V0, Uinit=U0, trajectory=True)
    ref = c_oracle.inmf(Es, k, lam, 6, W0, V0, Ps=Ps, Uinit=U0, trajectory=True)
    _check(out, ref, u=True)
    assert np.max(np.abs(out["traj"] - ref["traj"]) / ref["traj"]) < TOL_OBJ
    assert np.all(np.diff(out["traj"]) < 0)


@pytest.mark.parametrize("k", [24, 44])     # 44: the sweeps run once per half of the factors
def test_uinmf_dense_tensor_core_sweeps(k):
    """UINMF with count-structured shared AND unshared blocks: the stacked [E_i; P_i] sweeps run as dense-tile tcgen05
    contractions (rows = 1400 + 700: six gene blocks, 33 gene stages), checked against the FP64 oracle from identical
    inits; a different col_scale on an unshared block is refused."""
    import liger_b200
    from liger_b200 import synthetic as syn
    d, n, m, u = 2, 2300, 1400, 700
    Es, Ps, sc, psc = syn.make_numpy(d, n, m, 10, u=u, block=1150, seed=91, return_unshared_scales=True)
    lam = [5.0, 3.0]
    W0, V0, U0, _ = orc.seeded_init(1, m, k, [n] * d, [u] * d)
    Ec = [liger_b200.CountScaled(E, *s) for E, s in zip(Es, sc)]
    Pc = [liger_b200.CountScaled(P, *s) for P, s in zip(Ps, psc)]
    ctx = liger_b200.Context(Ec, k, lam, W0, V0, unsharedList=Pc, Uinit=U0)
    try:
        ctx.set_kernel_timing(True)
        ctx.iterate(2)
        names = set(ctx.kernel_times().keys())
        assert "k_dense_ltx" in names and "k_dense_xh" in names and "k_spmm_ltx" not in names
    finally:
        ctx.close()
    out = liger_b200.uinmf(Ec, Pc, k=k, lambda_=lam, niter=6, Winit=W0, Vinit=V0, Uinit=U0, trajectory=True)
    ref = c_oracle.inmf(Es, k, lam, 6, W0, V0, Ps=Ps, Uinit=U0, trajectory=True)
    _check(out, ref, u=True)
    assert np.max(np.abs(out["traj"] - ref["traj"]) / ref["traj"]) < TOL_OBJ
    gather = liger_b200.uinmf(Es, Ps, k=k, lambda_=lam, niter=6, Winit=W0, Vinit=V0, Uinit=U0)
    assert rel(out["W"], gather["W"]) < 1e-4 and rel(out["U"][1], gather["U"][1]) < 1e-4
    bad = [Pc[0], liger_b200.CountScaled(Ps[1], psc[1][0], psc[1][1] * 1.5)]
    with pytest.raises(Exception, match="col_scale differs|do not factor"):
        liger_b200.uinmf(Ec, bad, k=k, lambda_=lam, niter=1, Winit=W0, Vinit=V0, Uinit=U0)


def test_inmf_null_inits_draw_the_seeded_stream(pbmc):
    """RcppPlanc::inmf(..., Hinit = NULL, Vinit = NULL, Winit = NULL) (R/suggestK.R:78-88): the library draws W, V_i from R's
    Mersenne-Twister stream in .runINMF.list's order, so seed = 1 lands on the reference's golden factors."""
    import liger_b200
    g = pbmc["golden"]
    out = liger_b200.inmf(pbmc["Es"], k=20, lambda_=5.0, niter=30, Winit=None, Vinit=None, seed=1)
    rows = g["golden_rows"]
    assert rel(out["W"][rows], g["golden_W"]) < 5e-5
    for i, n in enumerate(pbmc["names"]):
        assert rel(out["H"][i].T, g[f"golden_H_{n}"]) < 5e-5
    other = liger_b200.inmf(pbmc["Es"], k=20, lambda_=5.0, niter=30, Winit=None, Vinit=None, seed=2)
    assert rel(other["W"], out["W"]) > 1e-3                      # another seed, another local optimum


def test_pageable_and_pinned_host_buffers_agree():
    """Inputs in pageable memory (what R hands over) go through the bounce ring of lgr_stage.cu in 4 MB chunks on four
    threads (here: 24 MB of values = 6 chunks), outputs come back the same way; page-locked buffers are copied directly.
    Same factors either way, and equal to the oracle's."""
    import liger_b200
    from liger_b200 import synthetic as syn
    from liger_b200.api import pinned_empty
    n, m, k = 30000, 2000, 30
    Es, _ = syn.make_numpy(1, n, m, k, block=5000, seed=3)
    E = Es[0]
    assert E.data.nbytes > 5 * (4 << 20)
    W0, V0, _, _ = orc.seeded_init(1, m, k, [n])
    a = liger_b200.inmf([E], k=k, lambda_=5.0, niter=2, Winit=W0, Vinit=V0)                       # pageable in and out
    pin = []
    for arr in (E.indptr, E.indices, E.data):
        h = pinned_empty(arr.shape, arr.dtype)
        h[...] = arr
        pin.append(h)
    b = liger_b200.inmf([(pin[0], pin[1], pin[2], E.shape)], k=k, lambda_=5.0, niter=2, Winit=W0, Vinit=V0,
                        pinned_outputs=True)
    assert rel(a["W"], b["W"]) < 1e-5 and rel(a["H"][0], b["H"][0]) < 1e-5 and rel(a["V"][0], b["V"][0]) < 1e-5
    ref = c_oracle.inmf([E], k, 5.0, 2, W0, V0)
    _check(a, ref)


# ----------------------------------------------------------------------------------------------
# dense-tile tensor-core sweeps (csrc/lgr_dense.cu): count-structured X = counts * row_scale * col_scale
# ----------------------------------------------------------------------------------------------
def _count_scaled(d, n, m, r, seed, block=1000):
    import liger_b200
    from liger_b200 import synthetic as syn
    Es, _, scales = syn.make_numpy(d, n, m, r, block=block, seed=seed, return_scales=True)
    return Es, [liger_b200.CountScaled(E, rs, cs) for E, (rs, cs) in zip(Es, scales)]


@pytest.mark.parametrize("name,d,n,m,k,niter", [
    ("C3-mini", 4, 3000, 2000, 30, 6),       # 32 gene stages, 12 cell tiles per dataset, 4 gene blocks
    ("ragged", 3, 1111, 333, 20, 5),         # n not a multiple of 256 / 32, m not a multiple of 64
    ("tiny", 2, 40, 70, 8, 4),               # less than one tile in every dimension
    ("wide", 1, 20000, 1100, 32, 3),         # more tiles than SMs x 1, k = 32 (full factor width), 3 gene blocks
    ("C4-mini", 2, 2500, 3000, 50, 5),       # k > 32: two passes over the entry streams (factor halves), cross-half Gram block
    ("k33", 3, 700, 300, 33, 4),             # one factor in the second half
    ("k62", 1, 1500, 260, 62, 3),            # the widest H solve the register kernels take
])
def test_dense_tensor_core_sweeps_match_oracle_and_gather(name, d, n, m, k, niter):
    """The tcgen05 sweeps (bf16 count tiles x 3-term bf16 split of the scaled factor, fp32 TMEM accumulators) against the
    FP64 C oracle from identical inits (north-star gates), and against the fp32 gather sweeps on the same input."""
    import liger_b200
    Es, Cs = _count_scaled(d, n, m, min(k, 10), seed=100 + d)
    W0, V0, _, _ = orc.seeded_init(1, m, k, [n] * d)
    out = liger_b200.inmf(Cs, k=k, lambda_=5.0, niter=niter, Winit=W0, Vinit=V0, trajectory=True)
    ref = c_oracle.inmf(Es, k, 5.0, niter, W0, V0, trajectory=True)
    _check(out, ref)
    assert np.max(np.abs(out["traj"] - ref["traj"]) / ref["traj"]) < TOL_OBJ
    gat = liger_b200.inmf(Cs, k=k, lambda_=5.0, niter=niter, Winit=W0, Vinit=V0, no_dense=True)
    assert rel(out["W"], gat["W"]) < 2e-5 and max(rel(a, b) for a, b in zip(out["H"], gat["H"])) < 2e-5
    assert out["timings"]["nnls_unconverged"] == 0


def test_dense_sweeps_large_counts_and_rejected_scales():
    """Counts up to 256 are exact in bf16; a count of 257, a non-integer ratio or wrong scales must be rejected with
    LGR_ERR_INVALID (the library verifies the structure entry by entry), never silently rounded."""
    import liger_b200
    from liger_b200._ffi import LigerB200Error
    rng = np.random.default_rng(11)
    m, n, k = 200, 900, 12
    N = sp.random(m, n, 0.08, format="csc", random_state=4, data_rvs=lambda s: rng.integers(1, 257, s).astype(float))
    N.data[:5] = 256.0
    rs = 1.0 / (0.5 + rng.random(m))
    cs = 1.0 / (100.0 + 50.0 * rng.random(n))
    X = sp.csc_matrix(N.multiply(rs[:, None]).multiply(cs[None, :]))
    W0, V0, _, _ = orc.seeded_init(2, m, k, [n])
    out = liger_b200.inmf([liger_b200.CountScaled(X, rs, cs)], k=k, lambda_=5.0, niter=4, Winit=W0, Vinit=V0, trajectory=True)
    ref = c_oracle.inmf([X], k, 5.0, 4, W0, V0, trajectory=True)
    _check(out, ref)
    assert np.max(np.abs(out["traj"] - ref["traj"]) / ref["traj"]) < TOL_OBJ
    bad = N.copy()
    bad.data[7] = 257.0
    Xb = sp.csc_matrix(bad.multiply(rs[:, None]).multiply(cs[None, :]))
    with pytest.raises(LigerB200Error) as e:
        liger_b200.inmf([liger_b200.CountScaled(Xb, rs, cs)], k=k, niter=1, Winit=W0, Vinit=V0)
    assert e.value.code == 1 and "integer counts" in str(e.value)
    with pytest.raises(LigerB200Error):
        liger_b200.inmf([liger_b200.CountScaled(X, rs * 1.37, cs)], k=k, niter=1, Winit=W0, Vinit=V0)
    # the same matrix without the structure runs through the gather sweeps
    out2 = liger_b200.inmf([Xb], k=k, lambda_=5.0, niter=2, Winit=W0, Vinit=V0)
    assert np.isfinite(out2["W"]).all()


def test_register_nnls_equals_shared_memory_nnls(pbmc):
    """The two H-solve kernels (register-resident, lgr_nnls_reg.cu; shared-memory factor, lgr_nnls.cu) solve the same
    strictly convex problems: same zero patterns away from degenerate coordinates, values to fp32 rounding."""
    import liger_b200
    W0, V0, _, _ = orc.seeded_init(1, 173, 20, [300, 300])
    a = liger_b200.inmf(pbmc["Es"], k=20, niter=7, Winit=W0, Vinit=V0)
    b = liger_b200.inmf(pbmc["Es"], k=20, niter=7, Winit=W0, Vinit=V0, legacy_nnls=True)
    for x, y in zip(a["H"], b["H"]):
        assert rel(x, y) < 1e-5
        scale = y.max()
        assert not np.any(((x == 0) & (y > 1e-5 * scale)) | ((y == 0) & (x > 1e-5 * scale)))
    assert rel(a["W"], b["W"]) < 1e-5


@pytest.mark.parametrize("k", [32, 31, 8])
def test_register_nnls_full_width_and_dead_factor(k):
    """k = 32 fills the padded factor row exactly (no spare padding column); a factor that is zero in W and every V_i
    makes the H-solve Gram singular, which must route that dataset's columns to the fallback kernel, not to NaNs."""
    import liger_b200
    from liger_b200 import synthetic as syn
    Es = [syn.make_numpy(1, n, 120, 8, block=500, seed=7 + i)[0][0] for i, n in enumerate((257, 64))]
    W0, V0, _, _ = orc.seeded_init(3, 120, k, [257, 64])
    out = liger_b200.inmf(Es, k=k, lambda_=5.0, niter=5, Winit=W0, Vinit=V0, trajectory=True)
    ref = c_oracle.inmf(Es, k, 5.0, 5, W0, V0, trajectory=True)
    _check(out, ref)
    W1 = W0.copy()
    V1 = [v.copy() for v in V0]
    W1[:, 2] = 0.0
    for v in V1:
        v[:, 2] = 0.0
    out = liger_b200.inmf(Es, k=k, lambda_=5.0, niter=1, Winit=W1, Vinit=V1)
    assert all(np.isfinite(h).all() for h in out["H"]) and np.isfinite(out["W"]).all()
    assert all(np.all(h[:, 2] == 0) for h in out["H"])          # the dead factor carries no loading after the H solve


def test_inmf_cold_start_equals_warm_start(pbmc):
    import liger_b200
    W0, V0, _, _ = orc.seeded_init(1, 173, 20, [300, 300])
    a = liger_b200.inmf(pbmc["Es"], k=20, niter=5, Winit=W0, Vinit=V0)
    b = liger_b200.inmf(pbmc["Es"], k=20, niter=5, Winit=W0, Vinit=V0, cold_start=True)
    assert rel(a["W"], b["W"]) < 1e-5 and rel(a["H"][0], b["H"][0]) < 1e-5


def test_inmf_niter0_uses_hinit_and_empty_cells(pbmc):
    """niter = 0 returns the inits and their objective; a cell without counts gets an all-zero H column."""
    import liger_b200
    W0, V0, _, H0 = orc.seeded_init(1, 173, 20, [300, 300])
    out = liger_b200.inmf(pbmc["Es"], k=20, niter=0, Winit=W0, Vinit=V0, Hinit=H0)
    want = sum(orc.objerr_i(H0[i].T, W0, V0[i], pbmc["Es"][i], 5.0) for i in range(2))
    assert abs(out["objErr"] - want) < 1e-5 * want
    assert rel(out["W"], W0) < 1e-15
    E = pbmc["Es"][0].tolil()
    E[:, 17] = 0
    E = sp.csc_matrix(E)
    E.eliminate_zeros()
    out = liger_b200.inmf([E, pbmc["Es"][1]], k=20, niter=3, Winit=W0, Vinit=V0)
    ref = orc.inmf([E, pbmc["Es"][1]], 20, 5.0, 3, W0, V0)
    assert np.all(out["H"][0][17] == 0)
    _check(out, ref)


def test_restarts_and_explicit_inits(pbmc):
    """nRandomStarts keeps restart 1's inits (`%||%`, R/integration.R:413-428); explicit inits are honoured
    (the optimizeNew* callers, R/optimizeNewParam.R:151-161)."""
    import liger_b200
    a = liger_b200.run_inmf_list(pbmc["Es"], k=12, nIteration=3, nRandomStarts=1, seed=4)
    b = liger_b200.run_inmf_list(pbmc["Es"], k=12, nIteration=3, nRandomStarts=3, seed=4)
    # the GPU run is not bitwise reproducible (fp32 atomics): equal-init restarts tie to ~1e-9, and a tie keeps the first
    # restart, as the reference's strict `<` does
    assert rel(a["W"], b["W"]) < 1e-5 and b["bestSeed"] == 4
    assert abs(a["objErr"] - b["objErr"]) < 1e-6 * a["objErr"]
    W0, V0, _, _ = orc.seeded_init(9, 173, 12, [300, 300])
    c = liger_b200.run_inmf_list(pbmc["Es"], k=12, nIteration=3, WInit=W0, VInit=V0)
    ref = orc.inmf(pbmc["Es"], 12, 5.0, 3, W0, V0)
    assert rel(c["W"], ref["W"]) < TOL_FACTOR


# ----------------------------------------------------------------------------------------------
# uinmf  (RcppPlanc::uinmf; R/integration.R:1285-1287) -- parity unpinned in the reference; vs the oracle
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k", [30, 44])     # 44: two-line factor rows, the k > 32 solver kernels on stacked [E; P] rows
def test_uinmf_matches_oracle(pbmc, k):
    import liger_b200
    Es = [E[:150] for E in pbmc["Es"]]
    Ps = [pbmc["Es"][0][150:], None]
    W0, V0, U0, _ = orc.seeded_init(3, 150, k, [300, 300], [23, 0])
    out = liger_b200.uinmf(Es, Ps, k=k, lambda_=[10.0, 1.0], niter=8, Winit=W0, Vinit=V0, Uinit=U0, trajectory=True)
    ref = orc.uinmf(Es, Ps, k, [10.0, 1.0], 8, W0, V0, U0, trajectory=True)
    _check(out, ref, u=True)
    assert out["U"][0].shape == (23, k) and out["U"][1] is None
    assert np.max(np.abs(out["traj"] - np.array(ref["traj"])) / np.array(ref["traj"])) < TOL_OBJ
    assert np.all(np.diff(out["traj"]) < 0)


def test_run_uinmf_list_like_reference_test(pbmc):
    """tests/testthat/test_factorization.R:79-98: k larger than the number of shared features is allowed."""
    import liger_b200
    Es = {n: E[:22] for n, E in zip(pbmc["names"], pbmc["Es"])}
    Ps = {"ctrl": pbmc["Es"][0][22:32], "stim": None}
    out = liger_b200.run_uinmf_list(Es, Ps, k=30, nIteration=2, nRandomStarts=2)
    assert out["W"].shape == (22, 30) and out["H"]["ctrl"].shape == (30, 300)
    assert list(out["U"].keys()) == ["ctrl"] and out["U"]["ctrl"].shape == (10, 30)


# ----------------------------------------------------------------------------------------------
# resident context + size-independent properties at a larger size
# ----------------------------------------------------------------------------------------------
def test_context_api_and_properties():
    import liger_b200
    from liger_b200 import synthetic as syn
    d, n, m, k, lam = 3, 6000, 800, 30, 5.0
    Es, _ = syn.make_numpy(d, n, m, 10, block=2000, seed=7)
    W0, V0, _, _ = orc.seeded_init(1, m, k, [n] * d)
    ctx = liger_b200.Context(Es, k, lam, W0, V0)
    t1 = ctx.iterate(4, trajectory=True)
    t2 = ctx.iterate(4, trajectory=True)
    traj = np.concatenate([t1, t2])
    assert np.all(np.diff(traj) < 1e-7 * traj[0])                      # monotone objective
    assert abs(ctx.objective() - traj[-1]) < 1e-6 * traj[-1]
    res = ctx.download(want_passive=True)
    ctx.close()
    one = liger_b200.inmf(Es, k=k, lambda_=lam, niter=8, Winit=W0, Vinit=V0)
    assert rel(res["W"], one["W"]) < 1e-5                                # 4 + 4 iterations == 8 iterations
    # KKT of the last H solve is not checkable after the W update; check H >= 0, mask consistency, and that the
    # final objective equals the oracle's formula on the downloaded factors (fp64)
    for i in range(d):
        H = res["H"][i]
        assert H.min() >= 0
        mask = res["H_passive"][i]
        bits = ((mask[:, None] >> np.arange(k, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(bool)
        assert np.all(H[~bits] == 0)
    want = sum(orc.objerr_i(res["H"][i].T, res["W"], res["V"][i], Es[i], lam) for i in range(d))
    assert abs(traj[-1] - want) < 1e-5 * want
    # and vs the C oracle after the same 8 iterations
    ref = c_oracle.inmf(Es, k, lam, 8, W0, V0)
    assert rel(res["W"], ref["W"]) < TOL_FACTOR
    for i in range(d):
        assert rel(res["H"][i], ref["H"][i]) < TOL_FACTOR and rel(res["V"][i], ref["V"][i]) < TOL_FACTOR
    assert abs(traj[-1] - ref["objErr"]) < TOL_OBJ * ref["objErr"]


@pytest.mark.parametrize("fused", [True, False])
def test_tensor_core_u_path(pbmc, fused, monkeypatch):
    """LGR_FLAG_TENSOR_U: u = CtC^-1 b on tcgen05 tensor cores (3xTF32 split, TMEM accumulators) -- same factors.
    fused: the epilogue of k_spmm_ltx_tiled (B tile still in shared memory); else the stand-alone k_pb_tc."""
    import liger_b200
    from liger_b200 import synthetic as syn
    monkeypatch.setenv("LGR_FUSE_U", "1" if fused else "0")
    W0, V0, _, _ = orc.seeded_init(1, 173, 20, [300, 300])
    a = liger_b200.inmf(pbmc["Es"], k=20, niter=30, Winit=W0, Vinit=V0, tensor_u=True, trajectory=True)
    ref = orc.inmf(pbmc["Es"], 20, 5.0, 30, W0, V0, trajectory=True)
    _check(a, ref)
    assert np.max(np.abs(a["traj"] - np.array(ref["traj"])) / np.array(ref["traj"])) < TOL_OBJ
    assert rel(a["W"], ref["W"]) < 1e-4                      # 3xTF32 keeps ~2^-21 per product
    # k > 32 (two 128-byte lines per row, N = 64 MMA) and ragged tiles (n not a multiple of 128)
    Es, _ = syn.make_numpy(2, 333, 150, 8, block=500, seed=5)
    W0, V0, _, _ = orc.seeded_init(2, 150, 40, [333, 333])
    b = liger_b200.inmf(Es, k=40, lambda_=2.5, niter=6, Winit=W0, Vinit=V0, tensor_u=True)
    refb = c_oracle.inmf(Es, 40, 2.5, 6, W0, V0)
    _check(b, refb)


def test_float32_values_input(pbmc):
    import liger_b200
    W0, V0, _, _ = orc.seeded_init(1, 173, 20, [300, 300])
    tup = [(E.indptr, E.indices, E.data.astype(np.float32), E.shape) for E in pbmc["Es"]]
    ctx = liger_b200.Context(tup, 20, 5.0, W0, V0, keep_f32=True)
    ctx.iterate(5)
    a = ctx.download()
    ctx.close()
    b = liger_b200.inmf(pbmc["Es"], k=20, niter=5, Winit=W0, Vinit=V0)
    assert rel(a["W"], b["W"]) < 1e-5


# ----------------------------------------------------------------------------------------------
# optimizeNewK / optimizeNewLambda / optimizeNewData / optimizeSubset  (R/optimizeNewParam.R)
# ----------------------------------------------------------------------------------------------
def _fit_pair(pbmc, k=8, niter=3):
    import liger_b200
    from liger_b200 import integration
    names = ["ctrl", "stim"]
    data = dict(zip(names, pbmc["Es"]))
    fit = integration.runINMF(data, k=k, lambda_=5.0, nIteration=niter, seed=1)
    base = orc.run_inmf_list(pbmc["Es"], k=k, lam=5.0, niter=niter, seed=1)
    ofit = {"W": base["W"], "V": base["V"], "H": base["H"], "lam": 5.0}
    return names, data, fit, ofit


def _cmp_fit(out, ref, names, tol=TOL_FACTOR):
    assert rel(out["W"], ref["W"]) < tol
    for i, n in enumerate(names):
        assert rel(out["V"][n], ref["V"][i]) < tol
        assert rel(out["H"][n], ref["H"][i]) < tol
        assert pattern_mismatches(out["H"][n].T, ref["H"][i].T) == 0
    assert abs(out["objErr"] - ref["objErr"]) / ref["objErr"] < TOL_OBJ


@pytest.mark.parametrize("k_new", [11, 5, 8])
def test_optimize_new_k(pbmc, k_new):
    from liger_b200 import optimize
    from oracle import optimize_oracle as oo
    names, data, fit, ofit = _fit_pair(pbmc)
    out = optimize.optimizeNewK(data, fit, kNew=k_new, nIteration=2, seed=1)
    if k_new == 8:
        assert out is fit           # test_factorization.R:106-107
        return
    ref = oo.optimize_new_k(pbmc["Es"], ofit, k_new, 2, seed=1)
    assert out["W"].shape == (173, k_new) and out["uns"]["factorization"]["k"] == k_new
    _cmp_fit(out, ref, names)


def test_optimize_new_lambda_subset_and_data(pbmc):
    from liger_b200 import optimize
    from oracle import optimize_oracle as oo
    names, data, fit, ofit = _fit_pair(pbmc)
    msgs = []
    out = optimize.optimizeNewLambda(data, fit, lambdaNew=4.0, nIteration=2, seed=1, warn=msgs.append)
    assert msgs and "New lambda less than current lambda" in msgs[0]          # test_factorization.R:114-115
    _cmp_fit(out, oo.optimize_new_lambda(pbmc["Es"], ofit, 4.0, 2, seed=1), names)
    rng = np.random.default_rng(0)
    idx = {"ctrl": np.sort(rng.choice(300, 150, replace=False)), "stim": np.sort(rng.choice(300, 150, replace=False))}
    out = optimize.optimizeSubset(data, fit, idx, nIteration=2)
    assert sum(h.shape[1] for h in out["H"].values()) == 300                  # test_factorization.R:119
    _cmp_fit(out, oo.optimize_subset(pbmc["Es"], ofit, [idx["ctrl"], idx["stim"]], 2), names)
    # new data as a new dataset inheriting ctrl's V (merge = FALSE), and appended to ctrl (merge = TRUE)
    new = sp.csc_matrix(pbmc["Es"][0][:, ::2] * 1.5)
    out = optimize.optimizeNewData(data, fit, {"ctrl2": new}, ["ctrl"], merge=False, nIteration=2)
    ref = oo.optimize_new_data(pbmc["Es"], ofit, [new], [0], False, niter=2)
    _cmp_fit(out, ref, names + ["ctrl2"])
    merged = sp.hstack([pbmc["Es"][0], new]).tocsc()
    out = optimize.optimizeNewData(data, fit, {"ctrl2": merged}, ["ctrl"], merge=True,
                                   newCells={"ctrl2": np.arange(300, 450)}, nIteration=2)
    ref = oo.optimize_new_data(pbmc["Es"], ofit, [merged], [0], True, new_cells=[np.arange(300, 450)], niter=2)
    assert out["H"]["ctrl"].shape[1] == 450                                   # cells were added (cf. :127)
    _cmp_fit(out, ref, names)


# ----------------------------------------------------------------------------------------------
# preprocessing feeding the path (normalize / selectGenes / scaleNotCenter), on the GPU
# ----------------------------------------------------------------------------------------------
def test_preprocessing_known_answers_on_gpu(pbmc):
    """The reference's own known answers (tests/testthat/test_preprocessing.R:184,190,262-263): 173 (union) / 161
    (intersection) variable genes, scaleData(ctrl)[3,5] = 0.4693316, scaleData(stim)[7,9] = 2.360295 (tol 1e-6);
    and the regenerated scaleData equals the oracle's, which the golden iNMF factors were reproduced from."""
    from liger_b200 import preprocess as pp
    raws, genes, numis = pbmc["raws"], pbmc["genes"], pbmc["nUMI"]
    norms = [pp.normalize(r) for r in raws]
    for n, r in zip(norms, raws):
        assert rel(n.toarray(), orc.normalize(r).toarray()) < 1e-14
    feats, sel = pp.selectGenes(norms, genes, numis, combine="union")
    assert len(feats) == 173
    feats_i, _ = pp.selectGenes(norms, genes, numis, combine="intersection")
    assert len(feats_i) == 161
    scaled = [pp.scaleNotCenter(n, g, feats) for n, g in zip(norms, genes)]
    assert abs(scaled[0][2, 4] - 0.4693316) < 1e-6
    assert abs(scaled[1][6, 8] - 2.360295) < 1e-6
    for s, e in zip(scaled, pbmc["Es"]):
        assert s.shape == e.shape and (s != 0).nnz == (e != 0).nnz
        assert rel(s.toarray(), e.toarray()) < 1e-13


def test_preprocessing_edge_cases_on_gpu():
    """empty rows (rms 0 -> stay 0), empty columns, a column chunk with the full-matrix divisor, one-cell rows."""
    from liger_b200 import preprocess as pp
    rng = np.random.default_rng(5)
    x = sp.random(57, 211, 0.08, format="csc", random_state=3, data_rvs=lambda n: rng.integers(1, 9, n).astype(float))
    x = sp.csc_matrix(x)
    x[5, :] = 0
    x[:, 7] = 0
    x.eliminate_zeros()
    keep = np.asarray(x.sum(axis=0)).ravel() > 0
    n_gpu = pp.normalize(x[:, keep])
    assert rel(n_gpu.toarray(), orc.normalize(x[:, keep]).toarray()) < 1e-14
    m, v = pp.rowMeansVars(n_gpu)
    dense = n_gpu.toarray()
    assert np.allclose(m, dense.mean(axis=1), rtol=1e-13, atol=0)
    assert np.allclose(v, dense.var(axis=1, ddof=1), rtol=1e-11, atol=1e-300)
    # chunked use (src/data_processing.cpp:148-150): sums of per-chunk outputs with the full ncol
    half = n_gpu.shape[1] // 2
    v1 = pp.rowMeansVars(n_gpu[:, :half], ncol=n_gpu.shape[1], means=m)[1]
    v2 = pp.rowMeansVars(n_gpu[:, half:], ncol=n_gpu.shape[1], means=m)[1]
    ref1 = orc.row_vars_sparse(n_gpu[:, :half], m, n_gpu.shape[1]) + orc.row_vars_sparse(n_gpu[:, half:], m, n_gpu.shape[1])
    assert np.allclose(v1 + v2, ref1, rtol=1e-12)
    genes = [f"g{i}" for i in range(57)]
    feats = genes[3:40]
    s_gpu = pp.scaleNotCenter(n_gpu, genes, feats)
    s_ref = orc.scale_not_center(n_gpu, genes, feats)
    assert s_gpu.shape == s_ref.shape and rel(s_gpu.toarray(), s_ref.toarray()) < 1e-13
    assert s_gpu[2, :].nnz == 0                   # the empty gene (row 5 of x) stays all-zero
    with pytest.raises(ValueError):
        pp.scaleNotCenter(n_gpu, genes, ["nope"])


def test_resident_pipeline_counts_to_factors(pbmc):
    """raw counts -> normalize -> selectGenes -> scaleNotCenter -> runINMF with X resident in HBM from the first upload on
    (lgr_prep_* + device views): same 173 genes, same scaleData as the oracle's, and the reference's golden factors."""
    import liger_b200
    from liger_b200 import preprocess as pp
    pipe = pp.Pipeline(pbmc["raws"], pbmc["genes"], pbmc["nUMI"])
    feats, sel = pipe.selectGenes(combine="union")
    assert len(feats) == 173 and [len(x) for x in sel] == [168, 166]
    assert len(pipe.selectGenes(combine="intersection")[0]) == 161
    Es = pipe.scaleNotCenter(feats)
    for i, e in enumerate(pbmc["Es"]):
        host = pipe.scaledToHost(i)
        assert host.shape == e.shape and rel(host.toarray(), e.toarray()) < 1e-13
    assert abs(pipe.scaledToHost(0)[2, 4] - 0.4693316) < 1e-6 and abs(pipe.scaledToHost(1)[6, 8] - 2.360295) < 1e-6
    g = pbmc["golden"]
    out = liger_b200.run_inmf_list(Es, k=20, lambda_=5.0, nIteration=30, seed=1)
    rows = g["golden_rows"]
    assert rel(out["W"][rows], g["golden_W"]) < 5e-5
    for i, n in enumerate(pbmc["names"]):
        assert rel(out["H"][i], g[f"golden_H_{n}"]) < 5e-5
        assert pattern_mismatches(out["H"][i], g[f"golden_H_{n}"]) == 0
    # a view dies with the buffers behind it: re-scaling (here with fewer features) or closing the pipeline invalidates
    # the views taken before, and the factorisation refuses them instead of reading freed memory
    stale = Es
    fresh = pipe.scaleNotCenter(feats[:150])
    with pytest.raises(Exception, match="stale or foreign device view"):
        liger_b200.run_inmf_list(stale, k=5, lambda_=5.0, nIteration=1, seed=1)
    ok = liger_b200.run_inmf_list(fresh, k=5, lambda_=5.0, nIteration=1, seed=1)
    assert ok["W"].shape == (150, 5)
    pipe.close()
    with pytest.raises(Exception, match="stale or foreign device view"):
        liger_b200.run_inmf_list(fresh, k=5, lambda_=5.0, nIteration=1, seed=1)


def test_cinmf_block_solves_and_refinement(pbmc):
    """inmf_solveH / solveV / solveW (R/cINMF.R:477-503) and runCINMF's fixed-W refinement loop (:392-407) as block solves
    on a resident context (the right-hand sides come from the library's fp32 sweeps over X, so the tolerance is the
    factor tolerance of the path, not FP64 round-off), against the oracle's restatement of the same three solves."""
    from liger_b200 import cinmf
    Es = pbmc["Es"]
    W, Vs, _, _ = orc.seeded_init(4, 173, 12, [300, 300])
    H0 = cinmf.inmf_solveH(None, W, Vs[0], Es[0], 5.0)
    assert rel(H0, orc.solve_H(W, Vs[0], Es[0], 5.0)) < 2e-5
    V0 = cinmf.inmf_solveV(H0, W, None, Es[0], 5.0)
    assert rel(V0, orc.solve_V(H0, W, Es[0], 5.0)) < 2e-5
    Hs = [H0, cinmf.inmf_solveH(None, W, Vs[1], Es[1], 5.0)]
    Wn = cinmf.inmf_solveW(Hs, W, [V0, Vs[1]], Es, 5.0)
    assert rel(Wn, orc.solve_W(Hs, [V0, Vs[1]], Es)) < 2e-5
    out = cinmf.refine_fixed_W(Es, W, Vs, lambda_=5.0, nIteration=2, seed=1)
    Vr = [v.copy() for v in Vs]
    Hr = [None, None]
    for _ in range(4):
        for i in range(2):
            Hr[i] = orc.solve_H(W, Vr[i], Es[i], 5.0)
        for i in range(2):
            Vr[i] = orc.solve_V(Hr[i], W, Es[i], 5.0)
    ref_obj = sum(orc.objerr_i(Hr[i].T, W, Vr[i], Es[i], 5.0) for i in range(2))
    for i in range(2):
        assert rel(out["H"][i], Hr[i]) < 1e-4 and rel(out["V"][i], Vr[i]) < 1e-4
    assert abs(out["objErr"] - ref_obj) < 1e-5 * ref_obj


def test_resident_block_solver_products_and_objective(pbmc):
    """lgr_ctx_solve / lgr_ctx_set_factors / lgr_ctx_products / lgr_ctx_objerr_i: one upload of X, then any sequence of
    block solves, raw products and per-dataset objectives from the resident state; checked against the oracle's block
    solves, scipy products and objerr_i."""
    from liger_b200 import cinmf
    Es = pbmc["Es"]
    k = 12
    W, Vs, _, _ = orc.seeded_init(7, 173, k, [300, 300])
    s = cinmf.BlockSolver(Es, k, 5.0, W, Vs)
    try:
        with pytest.raises(Exception, match="need an H"):
            s.ctx.solve("V")
        ltx = s.ctx.products("ltx")
        for i in range(2):
            assert rel(ltx[i], np.asarray((Es[i].T @ (W + Vs[i])))) < 2e-6
        Hs = s.solveH()
        Hr = [orc.solve_H(W, Vs[i], Es[i], 5.0) for i in range(2)]
        for i in range(2):
            assert rel(Hs[i], Hr[i]) < 2e-5
        xh = s.ctx.products("xh")
        for i in range(2):
            assert rel(xh[i], np.asarray(Es[i] @ Hs[i])) < 2e-6
        for i in range(2):      # objErr_i per dataset, and their sum
            assert abs(s.objErr(i) - orc.objerr_i(Hs[i].T, W, Vs[i], Es[i], 5.0)) < 1e-5 * s.objErr(i)
        assert abs(s.objErr() - (s.objErr(0) + s.objErr(1))) < 1e-9 * s.objErr()
        Vn = s.solveV()         # with the old W
        for i in range(2):
            assert rel(Vn[i], orc.solve_V(Hr[i], W, Es[i], 5.0)) < 5e-5
        Wn = s.solveW()         # with the new V
        assert rel(Wn, orc.solve_W(Hr, [orc.solve_V(Hr[i], W, Es[i], 5.0) for i in range(2)], Es)) < 5e-5
        # replacing H from the host invalidates the statistics: the next product / solve uses the new H
        H2 = [np.abs(h[::-1]).copy() for h in Hr]
        s.set(Hs=H2)
        xh2 = s.ctx.products("xh")
        assert rel(xh2[0], np.asarray(Es[0] @ H2[0])) < 2e-6
    finally:
        s.close()


# ----------------------------------------------------------------------------------------------
# quantileNorm on the device (f3; R/integration.R:1621-1711, src/quantile_norm.cpp)
# ----------------------------------------------------------------------------------------------
def test_quantile_norm_reproduces_reference_h_norm(pbmc):
    """lgr_quantile_norm on the reference's golden H with the defaults of its command log: the H.norm stored in
    data/pbmcPlot.rda to round-off (FP64 path, tolerance 1e-12), and the oracle's cluster labels exactly."""
    from liger_b200 import align
    from oracle import quantile_norm_oracle as qn
    g = pbmc["golden"]
    HList = {"ctrl": np.asarray(g["golden_H_ctrl"]), "stim": np.asarray(g["golden_H_stim"])}      # k x n
    res = align.quantile_norm_hlist(HList, reference="ctrl", return_info=True)
    gold = np.asarray(g["golden_Hnorm"])
    assert res["H.norm"].shape == gold.shape
    assert np.abs(res["H.norm"] - gold).max() < 1e-12
    Hs = [HList["ctrl"].T.copy(), HList["stim"].T.copy()]
    knn = [qn.knn_ann(h, 20, 0.9) for h in Hs]
    _, clusters = qn.quantile_norm(Hs, reference=0, knn=knn)
    assert (res["clusters"]["ctrl"] == clusters[0]).all() and (res["clusters"]["stim"] == clusters[1]).all()
    assert res["info"]["vote_sweeps"] >= 2 and res["info"]["sampled"] == 0
    # the front end: fit in, fit out; reference validation as in R
    fit = {"H": HList, "W": np.zeros((3, 20)), "uns": {}}
    out = align.quantileNorm(fit, reference="ctrl")
    assert np.abs(out["H.norm"] - gold).max() < 1e-12 and out["uns"]["alignmentMethod"] == "quantileNorm"
    with pytest.raises(ValueError, match="one existing dataset"):
        align.quantileNorm(fit, reference="nope")
    with pytest.raises(ValueError, match="one existing dataset"):
        align.quantileNorm(fit, reference=[True, True])


@pytest.mark.parametrize("case", ["sampled", "norefine_center_dims", "ref_last"])
def test_quantile_norm_matches_oracle(case):
    """Synthetic H with planted clusters: clusters above maxSample (R's sample() stream is replayed), refineKNN = FALSE with
    centring and useDims, and a reference that is not the first dataset (earlier datasets see the unwarped reference)."""
    from liger_b200 import align
    from oracle import quantile_norm_oracle as qn
    rng = np.random.default_rng(11)
    k, sizes = 6, [700, 500, 600]
    Hs = []
    for n in sizes:
        z = rng.integers(0, k, n)
        h = rng.gamma(0.4, 0.3, (n, k)).astype(np.float32).astype(np.float64)
        h[np.arange(n), z] += rng.gamma(3.0, 1.0, n)
        h[rng.random((n, k)) < 0.15] = 0.0
        h[5] = 0.0                                    # a cell with no positive loading: label -1
        Hs.append(h)
    kw = dict(quantiles=50, min_cells=20, n_neighbors=20, max_sample=1000, seed=1)
    pkw = dict(quantiles=50, minCells=20, nNeighbors=20, maxSample=1000, seed=1)
    ref = 0
    if case == "sampled":
        kw.update(max_sample=60, seed=7)
        pkw.update(maxSample=60, seed=7)
    elif case == "norefine_center_dims":
        kw.update(refine_knn=False, center=True, dims_use=[1, 2, 4, 6])
        pkw.update(refineKNN=False, center=True, useDims=[1, 2, 4, 6])
    else:
        ref = 2
    knn = None if not kw.get("refine_knn", True) else [qn.knn_ann(h, 20, 0.9) for h in Hs]
    out, clusters = qn.quantile_norm(Hs, reference=ref, knn=knn, **kw)
    res = align.quantile_norm_hlist([h.T for h in Hs], reference=ref + 1, return_info=True, **pkw)
    for i, name in enumerate(["1", "2", "3"]):
        assert (res["clusters"][name] == clusters[i]).all()
    assert np.abs(res["H.norm"] - np.vstack(out)).max() < 1e-11
    assert res["info"]["sampled"] == (1 if case == "sampled" else 0)
    assert res["info"]["warped_clusters"] > 0


def test_quantile_norm_on_resident_h(pbmc):
    """lgr_ctx_quantile_norm consumes the H that is still in HBM after the factorisation: same result as downloading H and
    calling lgr_quantile_norm on it."""
    import liger_b200
    from liger_b200 import align
    W, Vs, _, _ = orc.seeded_init(1, 173, 20, [300, 300])
    ctx = liger_b200.Context(pbmc["Es"], 20, 5.0, W, Vs)
    try:
        ctx.iterate(5)
        H = ctx.download()["H"]
        a = align.quantile_norm_resident(ctx, reference=0)
        b = align.quantile_norm_hlist([h.T for h in H], reference=1)
        assert np.abs(a["H.norm"] - b["H.norm"]).max() == 0.0
        assert all((x == y).all() for x, y in zip(a["clusters"], b["clusters"].values()))
        assert np.abs(a["H.norm"] - np.vstack(H)).max() > 1e-3        # something was warped
    finally:
        ctx.close()


# ----------------------------------------------------------------------------------------------
# centroidAlign on the device (f3; R/integration.R:2015-2084, src/centoidAlign.cpp) -- parity unpinned in the reference
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["defaults", "cluster_scaled_shift", "dims_lambda"])
def test_centroid_align_matches_oracle(pbmc, case):
    from liger_b200 import align
    from oracle import centroid_align_oracle as ca
    g = pbmc["golden"]
    HList = {"ctrl": np.asarray(g["golden_H_ctrl"]), "stim": np.asarray(g["golden_H_stim"])[:, :250]}
    Hs = [h.T for h in HList.values()]
    kw, okw = {}, {}
    if case == "cluster_scaled_shift":
        kw = dict(scaleCluster=True, centerCluster=True, shift=True)
        okw = dict(scale_cluster=True, center_cluster=True, shift=True)
    elif case == "dims_lambda":
        kw = dict(useDims=[2, 3, 5, 7, 11, 13], lambda_=[0.5, 3.0], centerEmb=False)
        okw = dict(dims_use=[2, 3, 5, 7, 11, 13], lam=[0.5, 3.0], center_emb=False)
    ref = ca.centroid_align(Hs, diagnosis=True, **okw)
    res = align.centroid_align_hlist(HList, diagnosis=True, **kw)
    assert res["aligned"].shape == ref["aligned"].shape
    assert rel(res["aligned"], ref["aligned"]) < 1e-11
    for key in ("raw_which.max", "Z_which.max", "R_which.max"):
        assert (res[key] == ref[key]).mean() > 0.999       # arg-max ties aside
    fit = align.alignFactors({"H": HList, "W": np.zeros((3, 20)), "uns": {}}, method="centroid", **kw)
    assert fit["uns"]["alignmentMethod"] == "centroidAlign" and rel(fit["H.norm"], ref["aligned"]) < 1e-11


def test_centroid_align_errors_and_resident(pbmc):
    import liger_b200
    from liger_b200 import align
    g = pbmc["golden"]
    HList = {"ctrl": np.asarray(g["golden_H_ctrl"]), "stim": np.asarray(g["golden_H_stim"])}
    # tests/testthat/test_factorization.R:221-222
    with pytest.raises(Exception, match="Negative values found prior to normalizing"):
        align.centroidAlign(HList, centerCluster=True, shift=False)
    with pytest.raises(ValueError, match="length 1 or 2"):
        align.centroidAlign(HList, lambda_=[1.0, 2.0, 3.0])
    W, Vs, _, _ = orc.seeded_init(1, 173, 20, [300, 300])
    ctx = liger_b200.Context(pbmc["Es"], 20, 5.0, W, Vs)
    try:
        ctx.iterate(5)
        H = ctx.download()["H"]
        a = align.centroid_align_resident(ctx)
        b = align.centroid_align_hlist([h.T for h in H])
        assert rel(a["aligned"], b["aligned"]) < 1e-13
    finally:
        ctx.close()


# ----------------------------------------------------------------------------------------------
# online iNMF, scenarios 1-3 (f4; R/integration.R:722-933 -> RcppPlanc::onlineINMF) -- parity unpinned in the reference
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k,mbs,epochs,hals", [(10, 100, 3, 1), (20, 250, 2, 2)])
def test_online_inmf_matches_oracle(pbmc, k, mbs, epochs, hals):
    """lgr_online_inmf against the numpy restatement of the published algorithm from the same seed: identical minibatch
    schedule and initialisation (R's RNG), factors within the path's tolerance, statistics A / B and objective too."""
    import liger_b200
    from oracle import online_oracle as oo
    Es = pbmc["Es"]
    ref = oo.online_inmf(Es, k, 5.0, max_epochs=epochs, minibatch_size=mbs, hals_iter=hals, seed=1)
    out = liger_b200.online_inmf(Es, k=k, lambda_=5.0, maxEpochs=epochs, minibatchSize=mbs, HALSiter=hals, seed=1)
    assert out["iterations"] == ref["iterations"] == 600 * epochs // mbs
    assert rel(out["W"], ref["W"]) < TOL_FACTOR
    for i in range(2):
        assert rel(out["V"][i], ref["V"][i]) < TOL_FACTOR
        assert rel(out["H"][i], ref["H"][i]) < TOL_FACTOR
        assert rel(out["A"][i], ref["A"][i]) < TOL_FACTOR and rel(out["B"][i], ref["B"][i]) < TOL_FACTOR
        assert out["H"][i].min() >= 0 and out["V"][i].min() > 0        # nonneg() floors W and V at 1e-16
    assert abs(out["objErr"] - ref["objErr"]) < TOL_OBJ * ref["objErr"]


def test_online_inmf_front_end_and_inits(pbmc):
    import liger_b200
    from liger_b200 import integration
    from oracle import online_oracle as oo
    Es = pbmc["Es"]
    data = dict(zip(["ctrl", "stim"], Es))
    W0, V0, _, _ = orc.seeded_init(3, 173, 8, [300, 300])
    ref = oo.online_inmf(Es, 8, 5.0, max_epochs=2, minibatch_size=120, seed=5, Winit=W0, Vinit=V0)
    fit = integration.runOnlineINMF(data, k=8, maxEpochs=2, minibatchSize=120, seed=5, WInit=W0, VInit=V0)
    assert fit["H"]["ctrl"].shape == (8, 300) and fit["A"]["stim"].shape == (8, 8) and fit["B"]["ctrl"].shape == (173, 8)
    assert rel(fit["W"], ref["W"]) < TOL_FACTOR and rel(fit["H"]["stim"].T, ref["H"][1]) < TOL_FACTOR
    with pytest.raises(ValueError, match="should be less than the number of cells"):
        liger_b200.online_inmf([Es[0][:, :5], Es[1]], k=8)
    with pytest.raises(ValueError, match="Cannot find complete online iNMF result"):
        integration.runOnlineINMF(data, k=8, newDatasets={"x": Es[0]})
    with pytest.raises(ValueError, match="overlap with existing datasets"):
        integration.runOnlineINMF(data, k=8, newDatasets={"ctrl": Es[0]}, WInit=W0, VInit=V0)


@pytest.mark.parametrize("k,mbs,epochs,hals", [(10, 60, 3, 1), (16, 100, 2, 2)])
def test_online_inmf_new_datasets_and_projection(pbmc, k, mbs, epochs, hals):
    """Scenarios 2 and 3 of runOnlineINMF (R/integration.R:755-812, 839-849): `stim` arrives after `ctrl` has been factorised.
    Scenario 2 learns W and the new dataset's V / A / B / H from the new cells only, with the old dataset's statistics in the
    W updates and its V / A / B untouched; scenario 3 only solves H against the given W.  Against the numpy restatement
    from the same state and seed."""
    import liger_b200
    from liger_b200 import integration
    from oracle import online_oracle as oo
    ctrl, stim = pbmc["Es"]
    first = oo.online_inmf([ctrl], k, 5.0, max_epochs=2, minibatch_size=75, seed=1)
    ref = oo.online_inmf_new_datasets(first["W"], first["V"], first["A"], first["B"], [stim], k, 5.0, max_epochs=epochs,
                                      minibatch_size=mbs, hals_iter=hals, seed=2)
    fit = integration.runOnlineINMF({"ctrl": ctrl}, k=k, newDatasets={"stim": stim}, maxEpochs=epochs, minibatchSize=mbs,
                                    HALSiter=hals, seed=2, WInit=first["W"], VInit={"ctrl": first["V"][0]},
                                    AInit={"ctrl": first["A"][0]}, BInit={"ctrl": first["B"][0]}, HInit={"ctrl": first["H"][0].T})
    assert rel(fit["W"], ref["W"]) < TOL_FACTOR
    assert np.array_equal(fit["V"]["ctrl"], first["V"][0]) and np.array_equal(fit["A"]["ctrl"], first["A"][0])
    assert np.array_equal(fit["B"]["ctrl"], first["B"][0]) and np.array_equal(fit["H"]["ctrl"], first["H"][0].T)
    assert rel(fit["V"]["stim"], ref["V"][1]) < TOL_FACTOR and rel(fit["H"]["stim"].T, ref["H"][0]) < TOL_FACTOR
    assert rel(fit["A"]["stim"], ref["A"][1]) < TOL_FACTOR and rel(fit["B"]["stim"], ref["B"][1]) < TOL_FACTOR
    assert abs(fit["objErr"] - ref["objErr"]) < TOL_OBJ * ref["objErr"]
    assert not np.allclose(fit["W"], first["W"])                      # W did move
    proj = integration.runOnlineINMF({"ctrl": ctrl}, k=k, newDatasets={"stim": stim}, projection=True, WInit=first["W"])
    assert set(proj.keys()) == {"H", "uns"} and proj["H"]["stim"].shape == (k, 300)
    assert rel(proj["H"]["stim"].T, oo.project(first["W"], [stim])[0]) < TOL_FACTOR


def test_online_inmf_from_column_chunks(pbmc):
    """lgr_online_inmf_chunked (the reference's on-disk path: H5SpMat chunks, R/integration.R:745-750): the matrices exist only
    behind a reader of column chunks; the result equals the in-memory call (same device copy, same V_i cells in double
    precision; up to the order of the atomic sums), for ragged last chunks, a reader that first asks for more room, and scenario 2; a failing reader raises."""
    import liger_b200
    Es = [E.tocsc() for E in pbmc["Es"]]
    calls = []

    def reader_of(mats):
        def read(i, c0, nc):
            calls.append((i, c0, nc))
            sub = mats[i][:, c0:c0 + nc]
            return sub.indptr, sub.indices, sub.data
        return read
    kw = dict(k=10, lambda_=5.0, maxEpochs=2, minibatchSize=120, seed=3)
    ref = liger_b200.online_inmf(Es, **kw)
    src = {"read": reader_of(Es), "shapes": [E.shape for E in Es], "nnz": [E.nnz for E in Es], "chunk_cols": 77, "max_chunk_nnz": 1}
    out = liger_b200.online_inmf([None, None], chunk_source=src, **kw)
    assert (0, 0, 77) in calls and (1, 231, 69) in calls and sum(1 for c in calls if c[2] == 1) == 20   # 3 x 77 + 69; 10 V cells each
    # (not bit for bit: the statistics B_i are summed by atomics, so two in-memory runs differ in the last bits as well)
    assert rel(out["W"], ref["W"]) < 1e-5 and abs(out["objErr"] - ref["objErr"]) < 1e-7 * ref["objErr"]
    for key in ("H", "V", "A", "B"):
        assert all(rel(x, y) < 1e-5 for x, y in zip(out[key], ref[key]))
    # scenario 2 through the reader: only the new dataset is read
    first = liger_b200.online_inmf([Es[0]], k=10, maxEpochs=2, minibatchSize=75, seed=1)
    st = dict(Winit=first["W"], Vinit=[first["V"][0]], Ainit=[first["A"][0]], Binit=[first["B"][0]])
    ref2 = liger_b200.online_inmf([Es[0]], newDatasets=[Es[1]], k=10, maxEpochs=2, minibatchSize=60, seed=2, **st)
    src2 = {"read": reader_of([Es[1]]), "shapes": [Es[1].shape], "nnz": [Es[1].nnz], "chunk_cols": 1000}
    out2 = liger_b200.online_inmf([None], newDatasets=[None], chunk_source=src2, k=10, maxEpochs=2, minibatchSize=60, seed=2, **st)
    assert rel(out2["W"], ref2["W"]) < 1e-5 and rel(out2["H"][1], ref2["H"][1]) < 1e-5

    def broken(i, c0, nc):
        raise IOError("disk gone")
    with pytest.raises(IOError, match="disk gone"):
        liger_b200.online_inmf([None, None], chunk_source=dict(src, read=broken), **kw)
    with pytest.raises(Exception, match="fewer non-zeros|more non-zeros"):
        liger_b200.online_inmf([None, None], chunk_source=dict(src, nnz=[Es[0].nnz + 5, Es[1].nnz]), **kw)


def test_alignment_and_online_fuzz():
    """Random small shapes and parameters (ragged sizes, k = 2.., quantiles = 1.., maxSample = 2.., eps = 0..2, single-dataset
    and non-first-reference cases) of quantileNorm / centroidAlign / online iNMF against their oracles (scripts/align_fuzz.py)."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "align_fuzz.py"), "2", "8"], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
