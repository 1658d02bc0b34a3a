"""CPU suite: the oracle against the reference's own golden vectors / known answers."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import rel
from oracle import inmf_oracle as orc
from oracle import c_oracle


def test_r_rng_known_answer():
    # set.seed(1); runif(3)  
    u = orc.RRandom(1).runif(3)
    assert np.allclose(u, [0.2655087, 0.3721239, 0.5728534], atol=5e-8)
    # the product's host RNG (C++) is the same stream
    import liger_b200
    v = liger_b200.RRandom(1).runif(1000, 0, 2)
    w = orc.RRandom(1).runif(1000, 0, 2)
    assert np.array_equal(v, w)
    assert np.array_equal(liger_b200.RRandom(12345).runif(700), orc.RRandom(12345).runif(700))


def test_preprocessing_known_answers(pbmc):
    g = pbmc["golden"]
    # tests/testthat/test_preprocessing.R:184,190,262-263
    assert int(g["n_union"]) == 173 and int(g["n_intersection"]) == 161
    assert abs(float(g["kat_ctrl_3_5"]) - 0.4693316) < 1e-6
    assert abs(float(g["kat_stim_7_9"]) - 2.360295) < 1e-6
    E = pbmc["Es"][0].toarray()
    assert abs(E[2, 4] - 0.4693316) < 1e-6
    assert pbmc["Es"][0].shape == (173, 300) and pbmc["Es"][1].shape == (173, 300)


def test_oracle_reproduces_reference_golden_factors(pbmc):
    """data/pbmcPlot.rda = runINMF(pbmc, k=20, lambda=5, nIteration=30, seed=1) (R/data.R:8-18)."""
    g = pbmc["golden"]
    out = orc.run_inmf_list(pbmc["Es"], k=20, lam=5.0, niter=30, seed=1, trajectory=True)
    rows = g["golden_rows"]
    assert rel(out["W"][rows], g["golden_W"]) < 1e-12
    for i, n in enumerate(pbmc["names"]):
        assert rel(out["V"][i][rows], g[f"golden_V_{n}"]) < 1e-12
        assert rel(out["H"][i], g[f"golden_H_{n}"]) < 1e-12
        assert np.array_equal(out["H"][i] == 0, g[f"golden_H_{n}"] == 0)
    traj = np.array(out["traj"])
    assert np.all(np.diff(traj) < 0)
    # objective trajectory recorded in SURVEY Appendix E
    assert np.allclose(traj[:5], [61343.051026, 43772.982015, 40131.427700, 38763.763611, 38062.001197], rtol=1e-9)
    assert abs(out["objErr"] - 36105.069129349) < 1e-6


def test_reference_test_config_k10_niter2(pbmc):
    # tests/testthat/test_factorization.R:49: runIntegration(pbmc, k = 10, nIteration = 2)
    out = orc.run_inmf_list(pbmc["Es"], k=10, lam=5.0, niter=2, seed=1, trajectory=True)
    assert np.allclose(out["traj"], [66551.145423, 49910.681818], rtol=1e-9)
    assert out["W"].shape == (173, 10) and out["H"][0].shape == (10, 300) and out["V"][1].shape == (173, 10)


def test_bmmc_trajectory(bmmc):
    W0, V0, _, _ = orc.seeded_init(1, 135, 20, [340, 340])
    out = orc.inmf(bmmc["Es"], 20, 5.0, 30, W0, V0, trajectory=True)
    t = np.array(out["traj"])
    assert np.allclose(t[[0, 1, 2, 29]], [76619.413, 67245.834, 63243.201, 58526.508], rtol=2e-8)
    assert np.all(np.diff(t) < 0)


def test_bpp_against_lawson_hanson():
    rng = np.random.default_rng(7)
    for k in (5, 20, 33):
        C = rng.random((3 * k, k))
        G = C.T @ C
        B = C.T @ (rng.random((3 * k, 40)) - 0.3)
        X = orc.bppnnls_prod(G, B)
        for j in range(B.shape[1]):
            x = orc.nnls_lawson_hanson(G, B[:, j])
            assert np.allclose(X[:, j], x, atol=1e-9)
        # KKT
        Y = G @ X - B
        assert X.min() >= 0 and Y[X == 0].min() > -1e-9 and np.abs(Y[X > 0]).max() < 1e-8


def test_c_oracle_matches_numpy_oracle(pbmc):
    W0, V0, _, _ = orc.seeded_init(1, 173, 20, [300, 300])
    a = c_oracle.inmf(pbmc["Es"], 20, 5.0, 8, W0, V0, trajectory=True)
    b = orc.inmf(pbmc["Es"], 20, 5.0, 8, W0, V0, trajectory=True)
    assert rel(a["W"], b["W"]) < 1e-11
    for x, y in zip(a["H"] + a["V"], b["H"] + b["V"]):
        assert rel(x, y) < 1e-11
    assert np.allclose(a["traj"], b["traj"], rtol=1e-12)
    rng = np.random.default_rng(3)
    C = rng.random((80, 30))
    G, B = C.T @ C, C.T @ rng.random((80, 257))
    assert rel(c_oracle.bppnnls_prod(G, B), orc.bppnnls_prod(G, B)) < 1e-12


def test_uinmf_oracle_monotone_and_stacked_equivalence(pbmc):
    Es = [E[:150] for E in pbmc["Es"]]
    Ps = [pbmc["Es"][0][150:], None]           # only the first dataset carries unshared features
    W0, V0, U0, _ = orc.seeded_init(3, 150, 20, [300, 300], [23, 0])
    b = orc.uinmf(Es, Ps, 20, [5.0, 2.0], 10, W0, V0, U0, trajectory=True)
    assert np.all(np.diff(b["traj"]) < 0)
    a = c_oracle.inmf(Es, 20, [5.0, 2.0], 10, W0, V0, Ps=Ps, Uinit=U0, trajectory=True)
    assert rel(a["W"], b["W"]) < 1e-11 and rel(a["U"][0], b["U"][0]) < 1e-11 and a["U"][1] is None
    assert abs(a["objErr"] - b["objErr"]) < 1e-7 * b["objErr"]


def test_objerr_formula_against_direct_norm(pbmc):
    rng = np.random.default_rng(1)
    E = pbmc["Es"][0]
    W, V, H = rng.random((173, 7)), rng.random((173, 7)), rng.random((7, 300))
    direct = np.linalg.norm(E.toarray() - (W + V) @ H) ** 2 + 3.5 * np.linalg.norm(V @ H) ** 2
    assert abs(orc.objerr_i(H, W, V, E, 3.5) - direct) < 1e-8 * direct


def test_synthetic_generator():
    from liger_b200 import synthetic as syn
    Es, Ps = syn.make_numpy(2, 1200, 400, 12, u=50, block=500)
    for E, P in zip(Es, Ps):
        assert E.shape == (400, 1200) and P.shape == (50, 1200)
        assert 0.03 < E.nnz / (400 * 1200) < 0.07
        assert E.data.min() > 0
        ss = np.asarray(E.multiply(E).sum(axis=1)).ravel() / (1200 - 1)
        assert np.allclose(ss[ss > 0], 1.0)


def test_optimize_oracle_invariants(pbmc):
    """oracle/optimize_oracle.py (R/optimizeNewParam.R): kNew == k is the identity, kNew < k keeps the factors with the
    largest loading norms and starts exactly from them (niter = 0), kNew > k starts from a point that does not lose
    to the old fit, a subset run starts from the subset's own H columns."""
    from oracle import optimize_oracle as oo
    Es = pbmc["Es"]
    base = orc.run_inmf_list(Es, k=8, lam=5.0, niter=3, seed=1)
    fit = {"W": base["W"], "V": base["V"], "H": base["H"], "lam": 5.0}
    assert oo.optimize_new_k(Es, fit, 8, 2) is fit
    down = oo.optimize_new_k(Es, fit, 5, 0)
    assert down["W"].shape == (173, 5) and down["H"][0].shape == (5, 300)
    cols = [int(np.argmin(np.abs(fit["W"] - down["W"][:, [j]]).sum(0))) for j in range(5)]
    assert len(set(cols)) == 5 and np.allclose(fit["W"][:, cols], down["W"])
    up0 = oo.optimize_new_k(Es, fit, 11, 0)
    assert up0["W"].shape == (173, 11) and up0["objErr"] <= base["objErr"] * (1 + 1e-9)
    assert min(x.min() for x in up0["H"]) >= 0 and up0["W"].min() >= 0
    idx = [np.arange(0, 300, 3), np.arange(1, 300, 2)]
    sub = oo.optimize_subset(Es, fit, idx, 0)
    assert np.allclose(sub["H"][1], fit["H"][1][:, idx[1]])
    lam2 = oo.optimize_new_lambda(Es, fit, 7.5, 2)
    assert lam2["lam"] == 7.5 and lam2["W"].shape == (173, 8)


# ----------------------------------------------------------------------------------------------
# quantileNorm (f3): the restatement against the H.norm the reference ships in data/pbmcPlot.rda
# ----------------------------------------------------------------------------------------------
def _golden_hlist(pbmc):
    g = pbmc["golden"]
    return [np.asarray(g["golden_H_ctrl"]).T.copy(), np.asarray(g["golden_H_stim"]).T.copy()]   # cell x factor


def test_quantile_norm_oracle_reproduces_reference_h_norm(pbmc):
    """quantileNorm(pbmcPlot) with the defaults of its command log: every row of the stored H.norm to round-off.  This
    needs both reference quirks: ANN's approximate kd-tree search (RANN::nn2(eps = 0.9)) and the in-place cluster vote."""
    from oracle import quantile_norm_oracle as qn
    Hs = _golden_hlist(pbmc)
    gold = np.asarray(pbmc["golden"]["golden_Hnorm"])
    knn = [qn.knn_ann(h, 20, 0.9) for h in Hs]
    out, clusters = qn.quantile_norm(Hs, reference=0, knn=knn)
    assert np.abs(np.vstack(out) - gold).max() < 1e-12
    # the reference dataset maps onto itself; 146 stim cells are warped
    assert np.abs(out[0] - Hs[0]).max() < 1e-12
    assert int((np.abs(out[1] - Hs[1]).max(axis=1) > 1e-12).sum()) == 146
    # without the in-place update, or with exact neighbours, the warped rows differ: the quirks are load-bearing
    out2, _ = qn.quantile_norm(Hs, reference=0, knn=knn, vote_in_place=False)
    assert np.abs(np.vstack(out2) - gold).max() > 1e-3
    out3, _ = qn.quantile_norm(Hs, reference=0)          # exact kNN
    assert np.abs(np.vstack(out3) - gold).max() > 1e-3


def test_quantile_norm_oracle_pieces():
    from oracle import quantile_norm_oracle as qn
    # stats::quantile type 7 known answers: quantile(c(1, 2, 4, 8, 16), c(0, .1, .5, .9, 1)) = 1.0 1.4 4.0 12.8 16.0
    q = qn.quantile7(np.array([16.0, 1.0, 4.0, 2.0, 8.0]), np.array([0, 0.1, 0.5, 0.9, 1.0]))
    assert np.allclose(q, [1.0, 1.4, 4.0, 12.8, 16.0], rtol=0, atol=1e-12)
    # approxfun(c(0, 0, 1, 2), c(1, 3, 5, 9), rule = 2): ties collapse to the mean (0 -> 2); outside -> end values
    v = qn.approx_rule2(np.array([0.0, 0.0, 1.0, 2.0]), np.array([1.0, 3.0, 5.0, 9.0]), np.array([-1.0, 0.0, 0.5, 1.5, 2.0, 7.0]))
    assert np.allclose(v, [2.0, 2.0, 3.5, 7.0, 9.0, 9.0])
    # set.seed(1); sample.int(10) under R >= 3.6 (sample.kind = "Rejection") = 9 4 7 1 2 5 3 10 6 8
    idx = qn.r_sample_index(orc.RRandom(1), 10, 10) + 1
    assert idx.tolist() == [9, 4, 7, 1, 2, 5, 3, 10, 6, 8]
    # the ANN search with eps = 0 is the exact search
    rng = np.random.default_rng(3)
    X = rng.gamma(0.6, 1.0, (400, 6))
    a = qn.knn_ann(X, 8, 0.0)
    e = qn.knn_exact(X, 8)
    assert (a == e).all()
    # max_factor: first maximum of the scaled columns, -1 for an all-zero row
    H = np.array([[1.0, 2.0], [3.0, 0.5], [0.0, 0.0], [2.0, 2.0]])
    assert qn.max_factor(H).tolist() == [2, 1, -1, 2]


def test_online_oracle_scenarios(pbmc):
    """The online iNMF restatement (unpinned against RcppPlanc, see its header): scenario 1 lowers the objective with more
    epochs; scenario 2 leaves the already factorised dataset's V / A / B alone, moves W, and fits the new dataset better
    than a projection onto the old W does; scenario 3 is the plain non-negative projection onto W."""
    from oracle import online_oracle as oo
    ctrl, stim = pbmc["Es"]
    k = 8
    a = oo.online_inmf([ctrl, stim], k, 5.0, max_epochs=1, minibatch_size=100, seed=1)
    b = oo.online_inmf([ctrl, stim], k, 5.0, max_epochs=4, minibatch_size=100, seed=1)
    assert b["objErr"] < a["objErr"] and b["iterations"] == 4 * a["iterations"] == 24
    first = oo.online_inmf([ctrl], k, 5.0, max_epochs=2, minibatch_size=75, seed=1)
    keep = [x.copy() for x in (first["V"][0], first["A"][0], first["B"][0])]
    s2 = oo.online_inmf_new_datasets(first["W"], first["V"], first["A"], first["B"], [stim], k, 5.0, max_epochs=3,
                                     minibatch_size=60, seed=2)
    assert s2["iterations"] == 300 * 3 // 60 and len(s2["V"]) == 2 and len(s2["H"]) == 1
    assert all(np.array_equal(x, y) for x, y in zip(keep, (s2["V"][0], s2["A"][0], s2["B"][0])))
    assert not np.allclose(s2["W"], first["W"]) and s2["W"].min() > 0 and s2["V"][1].min() > 0
    Hp = oo.project(first["W"], [stim])[0]
    assert Hp.shape == (300, k) and Hp.min() >= 0
    resid = lambda W, V, H: float(np.linalg.norm(stim.toarray() - (W + V) @ H.T))
    assert resid(s2["W"], s2["V"][1], s2["H"][0]) < resid(first["W"], np.zeros_like(first["W"]), Hp)
    # the projection is the NNLS solution: no feasible perturbation of a cell's loadings lowers its residual
    j = 17
    x = stim[:, j].toarray().ravel()
    base = np.linalg.norm(x - first["W"] @ Hp[j])
    rng = np.random.default_rng(0)
    for _ in range(20):
        h = np.maximum(Hp[j] + 1e-3 * rng.standard_normal(k), 0.0)
        assert np.linalg.norm(x - first["W"] @ h) >= base - 1e-12
