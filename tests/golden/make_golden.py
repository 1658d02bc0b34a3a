"""Regenerates tests/golden/*.npz from the reference's own fixtures.

Run in the BUILD container only (needs /root/reference/data/*.rda):
    python tests/golden/make_golden.py
It reads the reference's bundled objects with oracle/rda.py, restates the default
preprocessing (oracle/inmf_oracle.py) to obtain the exact `scaleData` inputs, and
stores (i) those inputs and (ii) the factors the reference itself ships in
data/pbmcPlot.rda (the golden W/V/H of runINMF(pbmc, k=20, lambda=5, nIteration=30,
seed=1), R/data.R:8-18).  Nothing here calls the oracle's iNMF: the golden arrays
are the reference's, untouched.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import rda  # noqa: E402
from oracle import inmf_oracle as orc  # noqa: E402

REF = "/root/reference/data"


def load_liger(path, name):
    obj = rda.read_rda(path)[name]
    ds = obj["datasets"]
    names = ds.names()
    cm = obj["cellMeta"]["listData"]
    cmn = cm.names()
    nUMI = cm.value[cmn.index("nUMI")].value
    dsf = cm.value[cmn.index("dataset")]
    levels = list(dsf.attrs["levels"].value)
    codes = dsf.value
    out = {}
    for nm, d in zip(names, ds.value):
        out[nm] = dict(obj=d, nUMI=nUMI[codes == levels.index(nm) + 1])
    return obj, out


def csc_pack(prefix, mat, store):
    mat = mat.tocsc()
    mat.sort_indices()
    store[prefix + "_indptr"] = mat.indptr.astype(np.int32)
    store[prefix + "_indices"] = mat.indices.astype(np.int32)
    store[prefix + "_data"] = mat.data.astype(np.float64)
    store[prefix + "_shape"] = np.array(mat.shape, dtype=np.int64)


def main():
    import scipy.sparse as sp
    # ---------------- pbmc: inputs ----------------
    _, pb = load_liger(f"{REF}/pbmc.rda", "pbmc")
    raws, genes, numis = [], [], []
    for nm in ("ctrl", "stim"):
        p, i, x, dim, dn = rda.dgc_to_csc(pb[nm]["obj"]["rawData"])
        raws.append(sp.csc_matrix((x, i, p), shape=dim))
        genes.append(dn[0])
        numis.append(pb[nm]["nUMI"])
    raws_pbmc, genes_pbmc, numis_pbmc = raws, genes, numis
    feats, scaled, sel = orc.default_pipeline(raws, genes, numis, combine="union")
    feats_int, _, _ = orc.default_pipeline(raws, genes, numis, combine="intersection")
    print("pbmc: selected", [len(s) for s in sel], "union", len(feats), "intersection", len(feats_int))
    store = {"features": np.array(feats), "n_union": len(feats), "n_intersection": len(feats_int)}
    for nm, s in zip(("ctrl", "stim"), scaled):
        csc_pack(f"scale_{nm}", s, store)
    sc = scaled[0].toarray(), scaled[1].toarray()
    # known answers of tests/testthat/test_preprocessing.R:184,190,262-263
    print("KAT scaleData(ctrl)[3,5] =", sc[0][2, 4], " scaleData(stim)[7,9] =", sc[1][6, 8])
    store["kat_ctrl_3_5"] = sc[0][2, 4]
    store["kat_stim_7_9"] = sc[1][6, 8]

    # ---------------- pbmcPlot: golden factors (the reference's own output) ----------------
    pp, ppd = load_liger(f"{REF}/pbmcPlot.rda", "pbmcPlot")
    Wg, Wdn = rda.dense(pp["W"])
    store["golden_W"] = Wg
    rowmap = np.array([feats.index(g) for g in Wdn[0]], dtype=np.int64)
    store["golden_rows"] = rowmap
    for nm in ("ctrl", "stim"):
        Hg, _ = rda.dense(ppd[nm]["obj"]["H"])
        Vg, Vdn = rda.dense(ppd[nm]["obj"]["V"])
        assert Vdn[0] == Wdn[0]
        store[f"golden_H_{nm}"] = Hg
        store[f"golden_V_{nm}"] = Vg
    # H.norm = quantileNorm(pbmcPlot) with the defaults stored in its command log (quantiles 50, reference "ctrl", minCells
    # 20, nNeighbors 20, eps 0.9, refineKNN TRUE, maxSample 1000, seed 1; RANN 2.6.1): pins the alignment step (f3)
    Hn, Hndn = rda.dense(pp["H.norm"])
    store["golden_Hnorm"] = Hn
    fac = pp["uns"]["factorization"]
    store["golden_k"] = int(fac.value[0].value[0])
    store["golden_lambda"] = float(fac.value[1].value[0])
    np.savez_compressed(os.path.join(HERE, "pbmc_golden.npz"), **store)
    # raw counts + gene names + nUMI of pbmc: inputs of the preprocessing parity tests (normalize / selectGenes /
    # scaleNotCenter on the GPU against the known answers of tests/testthat/test_preprocessing.R:184,190,262-263)
    rawstore = {}
    for nm, r, g, u in zip(("ctrl", "stim"), raws_pbmc, genes_pbmc, numis_pbmc):
        csc_pack(f"raw_{nm}", r, rawstore)
        rawstore[f"genes_{nm}"] = np.array(g)
        rawstore[f"nUMI_{nm}"] = np.asarray(u, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "pbmc_raw.npz"), **rawstore)

    # ---------------- bmmc: inputs only (config C2) ----------------
    _, bm = load_liger(f"{REF}/bmmc.rda", "bmmc")
    raws, genes, numis = [], [], []
    for nm in ("rna", "atac"):
        p, i, x, dim, dn = rda.dgc_to_csc(bm[nm]["obj"]["rawData"])
        raws.append(sp.csc_matrix((x, i, p), shape=dim))
        genes.append(dn[0])
        numis.append(bm[nm]["nUMI"])
    feats, scaled, sel = orc.default_pipeline(raws, genes, numis, combine="union")
    print("bmmc: selected", [len(s) for s in sel], "union", len(feats), "nnz", [s.nnz for s in scaled])
    store = {"features": np.array(feats)}
    for nm, s in zip(("rna", "atac"), scaled):
        csc_pack(f"scale_{nm}", s, store)
    np.savez_compressed(os.path.join(HERE, "bmmc_inputs.npz"), **store)


if __name__ == "__main__":
    main()
