"""Generate the committed fixtures under tests/golden/ (run in the build container, where
/root/reference exists; the GPU box only ever reads the resulting .npz files).

  example_sce.npz      : the reference's bundled fixture (data/example_sce.rda, R/data.R:1-23):
                         logcounts (genes x cells, f64), batch / cellTypes labels, ctl mask =
                         rows whose Ensembl id is in segList_ensemblGeneID$mouse$mouse_scSEG
                         (BASELINE.json configs[0]).
  golden_example.npz   : oracle outputs on that fixture via the scRUVIII path (k=20, supervised
                         M = one-hot(cellTypes), the deterministic stand-in for scReplicate).
  golden_sim.npz       : oracle outputs on a seeded ruvSimulate-distribution matrix (fastRUVIII exact,
                         the shape of tests/testthat/test-fastRUVIII.R:3-6) and a pseudoRUVIII case.
  golden_select.npz    : oracle outputs on the bundled fixture for the rows added later: scRUVIII with several ruvK
                         (silhouette coefficients, F-scores, chosen k; R/scRUVIII.R:80-114) and scRUVg (R/scRUVg.R:30-85).
                         `python tests/golden/make_golden.py select` writes only this file.

Usage: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from rdata_reader import factor_labels, read_rda  # noqa: E402

from oracle import ruv_oracle as orc  # noqa: E402

REF = "/root/reference/data"


def make_select():
    d = np.load(os.path.join(HERE, "example_sce.npz"))
    lc, ct, batch, ctl = d["logcounts"], d["cellTypes"], d["batch"], d["ctl"]
    ks = [5, 10, 20]
    res = orc.scRUVIII(lc, ct, ctl, ks, batch, cell_type=ct)
    g = orc.scRUVg(lc.T, ctl, 10)
    np.savez_compressed(os.path.join(HERE, "golden_select.npz"), ks=np.array(ks), sil=res["sil"], f_score=res["f_score"],
                        optimal=res["optimal_ruvK"], newY5_colsum=res[5]["newY"].sum(axis=0), newY5_fro=np.linalg.norm(res[5]["newY"]),
                        ruvg_newY_rows=g["newY"][:16], ruvg_newY_fro=np.linalg.norm(g["newY"]), ruvg_absW_colsum=np.abs(g["W"]).sum(axis=0),
                        ruvg_alpha_rownorm=np.linalg.norm(g["alpha"], axis=1))
    print("golden_select.npz", os.path.getsize(os.path.join(HERE, "golden_select.npz")) // 1024, "KiB", "sil", res["sil"],
          "f", res["f_score"], "optimal", res["optimal_ruvK"])


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "select":
        return make_select()
    sce = read_rda(os.path.join(REF, "example_sce.rda"))["example_sce"].attrs()
    assays = sce["assays"].attrs()[".xData"].value[".->data"].attrs()["listData"]
    names = assays.attrs()["names"].value
    lc = assays.value[names.index("logcounts")]
    dim = lc.attrs()["dim"].value
    genes = lc.attrs()["dimnames"].value[0].value
    logcounts = lc.value.reshape(dim[1], dim[0]).T.copy()  # R column-major -> genes x cells
    cd = sce["colData"].attrs()["listData"]
    cdn = cd.attrs()["names"].value
    cell_types = np.array(factor_labels(cd.value[cdn.index("cellTypes")]))
    batch = np.array(factor_labels(cd.value[cdn.index("batch")]))
    seg = read_rda(os.path.join(REF, "segList_ensemblGeneID.rda"))["segList_ensemblGeneID"]
    mouse = seg.value[seg.attrs()["names"].value.index("mouse")]
    scseg = set(mouse.value[mouse.attrs()["names"].value.index("mouse_scSEG")].value)
    ctl = np.array([g in scseg for g in genes])
    print("example_sce", logcounts.shape, "ctl", ctl.sum(), dict(zip(*np.unique(batch, return_counts=True))))
    np.savez_compressed(os.path.join(HERE, "example_sce.npz"), logcounts=logcounts, batch=batch,
                        cellTypes=cell_types, ctl=ctl)

    res = orc.scRUVIII(logcounts, cell_types, ctl, 20, batch)
    r = res[20]
    np.savez_compressed(os.path.join(HERE, "golden_example.npz"), fullalpha=r["fullalpha"],
                        newY_rows=r["newY"][:16], newY_colsum=r["newY"].sum(axis=0),
                        newY_fro=np.linalg.norm(r["newY"]), stand_mean=res["stand_mean"], stand_sd=res["stand_sd"])

    rng = np.random.default_rng(12345)
    L = orc.ruv_simulate(m=400, n=500, nc=400, n_celltypes=3, n_batch=2, lam=0.1, rng=rng)
    out = orc.fastRUVIII(L["Y"], L["M"], L["ctl"], 20, return_info=True)
    # pseudoRUVIII: 3 batches x 4 types x 6 folds pseudobulks from 1200 cells x 300 genes
    P = orc.planted_factors(1200, 300, 120, 8, n_types=4, n_batch=3, seed=777)
    which = (np.random.default_rng(5).permutation(1200) % 6) + 1
    bulks, keys = [], []
    for b in range(3):
        sel = P["batch"] == b
        bulk, nm = orc.create_pseudobulk(P["Y"][sel].T, P["types"][sel], which[sel], k_fold=6)
        bulks.append(bulk)
        keys += [int(x.split("_")[0]) for x in nm]
    Yp = np.concatenate(bulks)
    pr = orc.pseudoRUVIII(P["Y"], Yp, np.array(keys), P["ctl"], 8)
    assert np.all(L["Y"] == np.round(L["Y"])) and L["Y"].max() < 32000
    np.savez_compressed(os.path.join(HERE, "golden_sim.npz"),
                        sim_Y=L["Y"].astype(np.int16), sim_M=L["M"].astype(np.int8), sim_ctl=L["ctl"],
                        sim_newY_rows=out["newY"][:16], sim_newY_colsum=out["newY"].sum(axis=0),
                        sim_newY_fro=np.linalg.norm(out["newY"]), sim_fullalpha=out["fullalpha"],
                        ps_seed=777, ps_Y_rows=P["Y"][:4], ps_which=which.astype(np.int8),
                        ps_Ypseudo_colsum=Yp.sum(axis=0), ps_keys=np.array(keys),
                        ps_newY_rows=pr["newY"][:16], ps_newY_colsum=pr["newY"].sum(axis=0),
                        ps_newY_fro=np.linalg.norm(pr["newY"]), ps_fullalpha=pr["fullalpha"])
    for f in ("example_sce.npz", "golden_example.npz", "golden_sim.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
    make_select()


if __name__ == "__main__":
    main()
