"""bench.py's contract on a small workload (--quick: development shapes, not a bench value): one JSON
line with the keys the driver reads, parity green, launches counted; the same under torchrun with two
ranks on boxes with two GPUs (per-rank parity, variability merged on the device)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline", "parity")


def _line(cmd, timeout=600):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, "bench.py must print exactly one JSON line, got %d" % len(lines)
    return json.loads(lines[0])


def _check(d, n):
    for k in KEYS:
        assert k in d, k
    assert d["metric"] == "annotation_x_cell_zscores_per_sec_50bg" and d["unit"] == "z-scores/s"
    assert d["n_gpus"] == n and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] > 0
    assert d["parity"]["pass"] is True and d["parity"]["ranks_checked"] == n
    assert d["parity"]["integer_sums"]["observed_bit_equal"] and d["parity"]["integer_sums"]["background_bit_equal"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and 0 < r["frac"] < 1.5 and r["achieved"] > 0 and r["peak"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    if n == 1:                                                   # the CPU baseline leg runs at N = 1 only
        assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1


def test_bench_line_single_gpu():
    d = _line([sys.executable, "bench.py", "--quick", "--steps", "3", "--warmup", "3"])
    _check(d, 1)


def test_bench_line_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    d = _line([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
               "127.0.0.1", "--master-port", "29533", "bench.py", "--gpus", "2", "--quick", "--steps", "3", "--warmup", "3"])
    _check(d, 2)
    assert d.get("variability_merged_on_device") is True


def test_reference_arm_line():
    d = _line([sys.executable, "bench.py", "--impl", "reference", "--quick", "--steps", "1", "--warmup", "0"])
    assert d["impl"] == "reference" and d["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference")
