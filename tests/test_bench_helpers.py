"""CPU checks of bench.py's bookkeeping and of the synthetic background generator both arms share."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _bench():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_background_matrix_matches_bias_and_intensity():
    from chromvar_b200 import synthetic
    d = synthetic.make_config("tiny")
    bg = synthetic.background_matrix(d["counts"], d["bias"], niterations=20, seed=5)
    P = d["counts"].shape[0]
    assert bg.shape == (P, 20) and bg.dtype == np.int32 and bg.flags.f_contiguous
    assert bg.min() >= 1 and bg.max() <= P
    assert np.array_equal(bg, synthetic.background_matrix(d["counts"], d["bias"], niterations=20, seed=5))
    assert not np.array_equal(bg, synthetic.background_matrix(d["counts"], d["bias"], niterations=20, seed=6))
    fpp = np.log10(np.asarray(d["counts"].sum(axis=1)).ravel())
    for it in (0, 7, 19):
        assert np.corrcoef(d["bias"], d["bias"][bg[:, it] - 1])[0, 1] > 0.9
        assert np.corrcoef(fpp, fpp[bg[:, it] - 1])[0, 1] > 0.9


def test_config_object_is_the_same_for_both_arms():
    b = _bench()
    from chromvar_b200 import synthetic
    d = synthetic.make_config("tiny")
    c1 = b.workload_config("tiny", d, 4, b.BG_KIND)
    c2 = b.workload_config("tiny", d, 4, b.BG_KIND)
    assert c1 == c2 and c1["cells_total"] == 4 * d["counts"].shape[1] and c1["workload"] == "tiny"
    assert set(c1) == {"workload", "peaks", "cells_per_gpu", "annotations", "niterations", "nnz_per_gpu", "nnz_annotations",
                       "cells_total", "generator", "background", "l2"}


def test_roofline_classification_and_parity_helpers():
    b = _bench()
    peaks = {"hbm_gbs": 6543.1, "bf16_tflops": 1637.3}
    st = dict(path=1, operand_bits=4, n_cell_tiles=1, planes=1, cell_tile=256, contract_flops=1.0e14, overflow_columns=8476,
              few_cells=0)
    P, N, K = 100_000, 10_000, 870
    r = b.roofline_of(st, 12.3, 15.9, 2.0 * 51 * K * P * N, peaks, "measured", 1965.0, "cfg2_scatac_motifs", shape=(P, N, K))
    assert r["bound"] == "tensor" and 0.7 < r["frac"] < 0.8 and r["peak"] > 9000       # issue ceiling, not a bf16-derived figure
    assert r["frac_vs_bf16_derived_burst"] > 1.0 and r["ops_per_compulsory_byte"] > r["ridge_ops_per_byte"]
    # one 256-cell unit streaming the whole stacked operand: HBM-bound
    st4 = dict(path=1, operand_bits=8, n_cell_tiles=1, planes=2, cell_tile=256, contract_flops=2.3e13, overflow_columns=0, few_cells=0)
    r4 = b.roofline_of(st4, 8.2, 28.4, 2.0 * 51 * 870 * 500_000 * 200, peaks, "measured", 1965.0, "cfg4_bulk", shape=(500_000, 200, 870))
    assert r4["bound"] == "hbm" and 0.7 < r4["frac"] < 0.95 and r4["tensor_view"]["bound"] == "tensor"
    # few-cells kernel: one launch, tensor view
    st5 = dict(st4, few_cells=1)
    r5 = b.roofline_of(st5, 9.8, 15.0, 2.0 * 51 * 870 * 500_000 * 200, peaks, "measured", 1965.0, "cfg4_bulk", shape=(500_000, 200, 870))
    assert r5["kernel"] == "k_gather_contract" and r5["launches_per_step"] == 1 and r5["bound"] == "tensor"
    # parity helper: NA masks, infinities, tolerance
    ref = np.array([[1.0, np.nan, np.inf, 1e-9]])
    assert b.rel_err(ref.copy(), ref) == (0.0, True)
    got = ref.copy()
    got[0, 0] += 2e-6
    e, ok = b.rel_err(got, ref)
    assert ok and 1e-6 < e < 1e-5
    got[0, 1] = 0.0
    assert b.rel_err(got, ref)[1] is False
    assert abs(b.issue_peak_tflops(1965.0, 4) - 9522) < 5 and abs(b.issue_peak_tflops(1965.0, 8) - 4761) < 3
