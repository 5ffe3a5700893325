"""CPU checks of the drop-in boundary: the C-ABI library builds, loads, and exports every
symbol include/chromvar_b200.h declares; without a GPU it fails loudly instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "chromvar_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cv_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    from chromvar_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libchromvar_b200.so does not export %s" % n
    assert sorted(_lib.EXPORTS) == names


def test_version_and_struct_size():
    from chromvar_b200 import _lib
    lib = _lib.load()
    assert lib.cv_version() >= 100
    assert ctypes.sizeof(_lib.StageStats) == 8 * 8 + 5 * 8 + 10 * 4      # cv_stage_stats: 8 doubles, 5 int64, 10 int32


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from chromvar_b200 import _lib
    with pytest.raises(_lib.ChromVARError) as ei:
        _lib.Handle(0)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_merge_variability_host_function():
    """cv_merge_variability is pure host code (Chan merge of per-shard Welford partials)."""
    from chromvar_b200 import _lib
    rng = np.random.default_rng(1)
    z = rng.normal(size=(5, 90))
    z[1, 3] = np.nan
    shards = np.array_split(np.arange(90), 4)
    parts = np.zeros((4, 5, 4))
    for s, cols in enumerate(shards):
        for k in range(5):
            v = z[k, cols]
            f = v[np.isfinite(v)]
            parts[s, k] = [f.size, f.mean() if f.size else 0.0, ((f - f.mean()) ** 2).sum() if f.size else 0.0,
                           v.size - f.size]
    sd = _lib.merge_variability(parts, na_rm=True)
    ref = np.array([np.std(r[np.isfinite(r)], ddof=1) for r in z])
    assert np.allclose(sd, ref, rtol=1e-12)
    sd2 = _lib.merge_variability(parts, na_rm=False)
    assert np.isnan(sd2[1]) and np.allclose(np.delete(sd2, 1), np.delete(ref, 1), rtol=1e-12)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under chromvar_b200/ may reference it."""
    pkg = os.path.join(ROOT, "chromvar_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("test-only oracle", ""), f
