"""Generate the committed golden fixtures from the reference's bundled data.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Outputs (committed, small):
  deviations_test.npz  <- inst/extdata/deviations_test + the inputs hard-coded in
                          tests/testthat/test_compute_deviations.R:7,17,27
  mini.npz             <- data/mini_counts.rda, data/mini_ix.rda, data/mini_dev.rda
                          (config 1 of BASELINE.json)
  example_counts.npz   <- data/example_counts.rda (29269 x 50 counts + bias)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.setrecursionlimit(100000)
from rdata_reader import read_rda, read_rds  # noqa: E402

REF = "/root/reference"


def _assays(se):
    a = se.attr["assays"].attr[".xData"].value[".->data"]
    ld = a.attr["listData"]
    names = ld.attr["names"].value if "names" in ld.attr else None
    return names, ld.value


def _sparse_slots(m):
    dim = m.attr["Dim"].value
    return dict(i=m.attr["i"].value.astype(np.int32), p=m.attr["p"].value.astype(np.int32),
                x=np.asarray(m.attr["x"].value), dim=dim.astype(np.int64))


def _dense(m):
    d = m.attr["dim"].value
    return np.asarray(m.value, dtype=np.float64).reshape((d[0], d[1]), order="F")


def _df_cols(df):
    ld = df.attr["listData"]
    return dict(zip(ld.attr["names"].value, [v.value for v in ld.value]))


def main():
    # ---- golden vector 1: deviations_test ------------------------------------
    t = read_rds(os.path.join(REF, "inst/extdata/deviations_test"))
    names = t.attr["names"].value
    got = {n: _dense(v) for n, v in zip(names, t.value)}
    # inputs restated from tests/testthat/test_compute_deviations.R:7,17,27
    mat = np.array([0, 1, 2, 1, 1, 0, 2, 0, 1, 0, 1, 1], dtype=np.float64).reshape((3, 4), order="F")
    bg = np.array([2, 2, 2, 1, 1, 3, 2, 2, 1], dtype=np.int32).reshape((3, 3))  # byrow=TRUE
    np.savez(os.path.join(HERE, "deviations_test.npz"),
             counts=mat, bg=bg,
             set0=np.array([1, 2, 3], np.int32), set1=np.array([1, 2], np.int32),
             set2=np.array([2, 3], np.int32),
             deviations=got["deviations"], z=got["z"])
    print("deviations_test:", got["deviations"].shape, got["z"].shape)

    # ---- golden vector 2: mini_* (config 1) -----------------------------------
    mc = read_rda(os.path.join(REF, "data/mini_counts.rda"))["mini_counts"]
    _, (cm,) = _assays(mc)
    c = _sparse_slots(cm)
    bias = np.asarray(_df_cols(mc.attr["rowRanges"].attr["elementMetadata"])["bias"], np.float64)
    coldata = _df_cols(mc.attr["colData"])
    mi = read_rda(os.path.join(REF, "data/mini_ix.rda"))["mini_ix"]
    an, (im,) = _assays(mi)
    a = _sparse_slots(im)
    md = read_rda(os.path.join(REF, "data/mini_dev.rda"))["mini_dev"]
    dn, dms = _assays(md)
    assert dn == ["deviations", "z"], dn
    rd = _df_cols(md.attr["elementMetadata"])
    np.savez_compressed(
        os.path.join(HERE, "mini.npz"),
        counts_i=c["i"], counts_p=c["p"], counts_x=c["x"].astype(np.float64), counts_dim=c["dim"],
        bias=bias, cell_type=np.array(coldata["Cell_Type"]), depth=np.asarray(coldata["depth"]),
        ix_i=a["i"], ix_p=a["p"], ix_x=a["x"].astype(np.int8), ix_dim=a["dim"],
        ix_assay=np.array(an), ix_names=np.array(_df_cols(mi.attr["colData"])["name"]),
        dev=_dense(dms[0]), z=_dense(dms[1]),
        dev_names=np.array(rd["name"]),
        fractionMatches=np.asarray(rd["fractionMatches"], np.float64),
        fractionBackgroundOverlap=np.asarray(rd["fractionBackgroundOverlap"], np.float64))
    print("mini_counts:", c["dim"], "nnz", len(c["i"]), "sum", c["x"].sum(), "bias[:3]", bias[:3])
    print("mini_ix:", a["dim"], a["p"], "mini_dev:", _dense(dms[0]).shape, rd["fractionMatches"])

    # ---- example_counts (bigger CPU parity case) -------------------------------
    ec = read_rda(os.path.join(REF, "data/example_counts.rda"))["example_counts"]
    _, (em,) = _assays(ec)
    e = _sparse_slots(em)
    ebias = _df_cols(ec.attr["rowRanges"].attr["elementMetadata"]).get("bias")
    xs = e["x"]
    assert np.all(xs == np.round(xs)) and xs.max() < 65536
    out = dict(counts_i=e["i"], counts_p=e["p"], counts_x=xs.astype(np.uint16), counts_dim=e["dim"])
    if ebias is not None:
        out["bias"] = np.asarray(ebias, np.float64)
    np.savez_compressed(os.path.join(HERE, "example_counts.npz"), **out)
    print("example_counts:", e["dim"], "nnz", len(e["i"]), "bias" if ebias is not None else "no bias")


if __name__ == "__main__":
    main()
