"""CPU checks of the .Call glue (rpkg/src/cv_rglue.c) compiled against the stand-in R runtime
(rpkg/rstub/): the shared object loads, R_init_chromVARb200 registers exactly the routines the R shim
calls with the arities it uses (src/RcppExports.cpp:121-137 is the reference's pattern), dynamic
lookup is switched off, and the R-level error paths that need no GPU work."""
import os
import re
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def R():
    from rcall import RSession
    return RSession()


def test_registered_table_matches_the_r_shim(R):
    from rcall import RSession  # noqa: F401
    expected = {"cvb_init": 2, "cvb_shutdown": 0, "cvb_n_devices": 0, "cvb_set_seed": 1, "cvb_set_option": 2,
                "cvb_set_counts": 4, "cvb_set_counts_dense": 1, "cvb_fragments": 2, "cvb_expectations": 4,
                "cvb_background_peaks": 4, "cvb_background_peaks_knn": 3, "cvb_deviations": 5,
                "cvb_deviations_csc": 8, "cvb_row_sds": 2, "cvb_row_sds_perm": 3}
    assert R.routines == expected
    assert R.dll.rstub_dynamic_symbols() == 0            # R_useDynamicSymbols(dll, FALSE)
    # every .Call in the R shim names a registered routine and passes that many arguments
    src = open(os.path.join(ROOT, "rpkg", "R", "chromvar_b200.R")).read()
    calls = re.findall(r'\.Call\("(\w+)"((?:[^()]|\((?:[^()]|\([^()]*\))*\))*)\)', src)
    assert calls
    seen = set()
    for name, rest in calls:
        assert name in expected, name
        depth, n_args, cur = 0, 0, ""
        for ch in rest:
            if ch in "([":
                depth += 1
            elif ch in ")]":
                depth -= 1
            if ch == "," and depth == 0:
                if cur.strip():
                    n_args += 1
                cur = ""
            else:
                cur += ch
        if cur.strip():
            n_args += 1
        n_args -= 1                                      # PACKAGE = "chromVARb200"
        assert "PACKAGE" in rest
        assert n_args == expected[name], (name, n_args, expected[name])
        seen.add(name)
    assert {"cvb_init", "cvb_set_counts", "cvb_deviations_csc", "cvb_deviations", "cvb_background_peaks",
            "cvb_expectations", "cvb_row_sds", "cvb_row_sds_perm"} <= seen


def test_lookup_and_arity_errors_like_r(R):
    from rcall import RError
    with pytest.raises(RError, match="not in DLL"):
        R.call("cvb_no_such_routine")
    with pytest.raises(RError, match="Incorrect number of arguments"):
        R.call("cvb_set_seed", 1.0, 2.0)


def test_calls_before_init_stop(R):
    from rcall import RError
    R.call("cvb_shutdown")
    assert R.call("cvb_n_devices")[0] == 0
    for name, args in (("cvb_set_seed", (1.0,)), ("cvb_fragments", (3, 4)), ("cvb_row_sds", ([[1.0]], False))):
        with pytest.raises(RError, match="cvb_init"):
            R.call(name, *args)
