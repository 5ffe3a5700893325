#!/usr/bin/env python
"""bench.py -- annotation x cell z-scores per second (50 background iterations).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference

A "step" is one pass of the hot path (compute_deviations_core: counts layout, annotation
preparation, gather-contraction with fused epilogue, variability merge, ABI layout) over one
batch of synthetic scATAC input: BASELINE.json config 2 (100k peaks x 10k cells, ~3 % nonzero,
870 motif-like annotations, 50 background iterations) per GPU -- weak scaling, cells shard.

  value : device-timed (CUDA events on the library's stream), inputs resident in HBM.
  e2e   : the same metric through the C ABI with pinned HOST buffers, every step paying the
          H2D of counts / annotations / background / expectation and the D2H of both K x N
          result matrices; e2e.pageable: the same call on pageable host memory (what R passes).
  parity: at every N, every rank checks a random annotation sample of its own results against the
          CPU restatement (oracle/), integer observed / background sums included.
  workloads: N = 1: configs 3, 4 and the config-5 shard (device-timed, parity, roofline, timed
          getBackgroundPeaks); N > 1: the atlas (one config-5 shard per GPU) and `single_process`,
          the same e2e call through cv_multi from ONE process driving all GPUs.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "annotation_x_cell_zscores_per_sec_50bg"
UNIT = "z-scores/s"
WORKLOAD = "cfg2_scatac_motifs"
NITER = 50


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks_json():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def ncu_traffic(kernel, workload, ms_per_launch=None):
    """dram bytes (read + write) per launch of `kernel` on `workload` from the committed ncu --set
    full capture (profiles/traffic.json, from the .ncu-rep files summarised under profiles/), or None.
    Where the capture is one cell chunk of a step, the figure is scaled to the step's average launch
    by duration (traffic is proportional to the cells of the chunk)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        v = t.get(workload, {}).get(kernel)
        det = t.get(workload + "_detail", {}).get(kernel)
        if v is not None and det and ms_per_launch:
            return int(v * ms_per_launch / det["captured_launch_ms"])
        return v
    except Exception:
        return None


class ClockSampler:
    """SM clock / throttle reasons sampled every 100 ms during the timed region.  NVML in-process
    (nvmlInit happens in the constructor, before any timed work: spawning nvidia-smi inside the timed
    region takes driver locks for tens of milliseconds and shows up as slow steps); nvidia-smi -lms
    only when the NVML binding is missing."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, mode="thread"):
        self.mode = mode                  # "thread": poll every 100 ms beside the run; "inline": sample() between steps; "off"
        self.index, self.proc, self.lines, self.samples = index, None, [], []
        self.nv = self.dev = None
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # honour CUDA_VISIBLE_DEVICES: `index` is the CUDA ordinal
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        masks = (("hw_slowdown", getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8)),
                 ("hw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40)),
                 ("sw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20)),
                 ("sw_power_cap", getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)))
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                self.samples.append((sm, [n for n, m in masks if r & m]))
            except Exception:
                pass
            self._stop.wait(0.1)

    def sample(self):
        """One sample taken by the caller's thread (inline mode: between timed steps)."""
        if self.nv is None:
            return
        nv = self.nv
        try:
            sm = float(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
            masks = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
            self.samples.append((sm, [n for n, m in masks if r & m]))
        except Exception:
            pass

    def start(self):
        if self.mode != "thread":
            return
        if self.nv is not None:
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.mode == "off":
            return None
        if self.nv is not None:
            if self.mode == "thread":
                self._stop.set()
                self.t.join(timeout=1)
            if not self.samples:
                return None
            reasons = set()
            for _, rs in self.samples:
                reasons.update(rs)
            return {"sm_mhz": float(np.median([x for x, _ in self.samples])), "sm_max_mhz": self.max_sm,
                    "samples": len(self.samples), "reasons": sorted(reasons), "source": "nvml"}
        if not self.proc:
            return None
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "samples": len(sm),
                "reasons": sorted(reasons), "source": "nvidia-smi"}


def make_workload(rank, cells, quick, workload=None, gen="host", device=None):
    """gen: "host" (numpy generator), "device" (same model, draws and duplicate merging on the GPU:
    make_counts_device), "auto" = device for configurations of >= 5e8 expected nonzeros (the atlas
    shard takes minutes on the host)."""
    from chromvar_b200 import synthetic
    t0 = time.time()
    if quick:
        d = synthetic.make_config("small", seed_offset=rank)
        name = "small"
    else:
        over = {"N": cells} if cells else {}
        name = workload or WORKLOAD
        cfg = dict(synthetic.CONFIGS[name])
        cfg.update(over)
        big = cfg["density"] < 1.0 and cfg["P"] * cfg["N"] * cfg["density"] >= 1e8
        use_dev = device is not None and (gen == "device" or (gen == "auto" and big))
        d = synthetic.make_config(name, seed_offset=rank, device=device if use_dev else None, **over)
        if use_dev:
            import torch
            torch.cuda.empty_cache()
    if not hasattr(d["counts"], "indptr"):
        import scipy.sparse as sp
        d["dense"] = d["counts"]                       # bulk (config 4): dense column-major counts
        d["counts"] = sp.csc_matrix(d["counts"])
    log("[bench] rank %d: synthetic %s %s nnz=%d (%.1fs)" % (rank, name, d["counts"].shape, d["counts"].nnz,
                                                            time.time() - t0))
    return d, name


def workload_config(name, d, ws, background):
    """The `config` object of the JSON line: what the workload IS.  Built by this one function for both
    arms so that their lines describe the same configuration key for key (engine choices such as
    kernel / tile / step call live under "engine", not here)."""
    C = d["counts"]
    return {"workload": name, "peaks": int(C.shape[0]), "cells_per_gpu": int(C.shape[1]),
            "annotations": int(len(d["annoptr"]) - 1), "niterations": NITER, "nnz_per_gpu": int(C.nnz),
            "nnz_annotations": int(d["annoptr"][-1]), "cells_total": int(C.shape[1]) * int(ws),
            "generator": d["cfg"].get("generator", "numpy (make_counts)"), "background": background,
            "l2": "256 MB flush write between timed steps"}


BG_KIND = "synthetic.background_matrix(seed 20173): numpy generative sampler, same matrix for both arms"


def sets_of(d, ks):
    ptr, idx = d["annoptr"], d["annoidx"]
    return [idx[ptr[k]:ptr[k + 1]].astype(np.int64) for k in ks]


# ------------------------------------------------------------------------------------------------
# parity: results against the CPU restatement of the reference (oracle/), random annotation sample,
# integer sums included (SURVEY 8d "parity assertions in every timed run")
# ------------------------------------------------------------------------------------------------
def rel_err(got, ref):
    """(max abs(d) / max(abs(ref), 1e-3), NA / NaN / inf masks equal)"""
    ng, nr = np.isnan(got), np.isnan(ref)
    ig, ir = np.isinf(got), np.isinf(ref)
    ok = bool(np.array_equal(ng, nr) and np.array_equal(ig, ir) and np.array_equal(got[ig], ref[ir]))
    m = ~(nr | ng | ir | ig)
    worst = float((np.abs(got[m] - ref[m]) / np.maximum(np.abs(ref[m]), 1e-3)).max()) if m.any() else 0.0
    return worst, ok


def parity_block(eng, d, bg, e, got_dev, got_z, rank, workers, n_anno, cell_lo, cell_hi, slab=384, timed=False):
    """Random sample of n_anno annotations: deviations / z of cells [cell_lo, cell_hi) of the results
    `got_*` (K x N, whatever call produced them) against the oracle; observed and background integer
    sums bit-equal on a slab of `slab` cells (a separate call of the same kernels with the sums kept).
    Leaves the slab resident on the engine: call it last.  Returns (report, seconds of the oracle run,
    annotations x cells it covered)."""
    from oracle import chromvar_oracle as orc
    C = d["counts"]
    P, N = C.shape
    K = len(d["annoptr"]) - 1
    rng = np.random.default_rng(9000 + rank)
    ks = np.sort(rng.choice(K, size=min(n_anno, K), replace=False))
    Csub = C[:, cell_lo:cell_hi].tocsc() if (cell_lo, cell_hi) != (0, N) else C
    farm = orc.TaskFarm(Csub, sets_of(d, range(K)), bg, e, workers)
    t0 = time.perf_counter()
    ref_z, ref_dev = farm.run(ks)
    dt = time.perf_counter() - t0
    farm.close()
    ez, okz = rel_err(np.asarray(got_z[ks][:, cell_lo:cell_hi]), ref_z)
    ed, okd = rel_err(np.asarray(got_dev[ks][:, cell_lo:cell_hi]), ref_dev)
    rep = {"annotations_checked": int(ks.size), "annotation_sample": "random, seed %d" % (9000 + rank),
           "cells": int(cell_hi - cell_lo), "tolerance": 1e-5, "max_rel_err": max(ez, ed),
           "na_masks_equal": bool(okz and okd)}
    # integer sums on a slab of cells
    c0 = int(rng.integers(0, max(1, N - slab + 1)))
    c1 = min(N, c0 + slab)
    Cs = C[:, c0:c1].tocsc()
    eng.set_option("check_zero_peaks", 0)
    eng.set_counts_csc(P, c1 - c0, Cs.indptr, Cs.indices, Cs.data)
    got = eng.deviations(d["annoptr"], d["annoidx"].astype(np.int32), bg, e, index_base=1, sums=True)
    farm = orc.TaskFarm(Cs, sets_of(d, range(K)), bg, e, workers)
    rz, rd, robs, rsamp = farm.run(ks, sums=True)
    farm.close()
    obs_eq = bool(np.array_equal(got["observed"][ks], robs.astype(np.int64)))
    samp_eq = bool(np.array_equal(got["sampled"][ks], rsamp.astype(np.int64)))
    e2, ok2 = rel_err(got["z"][ks], rz)
    e3, ok3 = rel_err(got["dev"][ks], rd)
    rep["integer_sums"] = {"cells": int(c1 - c0), "first_cell": c0, "observed_bit_equal": obs_eq,
                           "background_bit_equal": samp_eq, "sums_compared": int(ks.size * (NITER + 1) * (c1 - c0)),
                           "max_rel_err_slab": max(e2, e3)}
    rep["max_rel_err"] = max(rep["max_rel_err"], e2, e3)
    rep["na_masks_equal"] = bool(rep["na_masks_equal"] and ok2 and ok3)
    rep["pass"] = bool(rep["na_masks_equal"] and rep["max_rel_err"] <= 1e-5 and obs_eq and samp_eq)
    return rep, dt, int(ks.size) * int(cell_hi - cell_lo)


def reduce_parity(rep, ws, dist, dev):
    """All ranks -> one verdict: pass only if every rank passed; the worst error."""
    if ws == 1:
        rep["ranks_checked"] = 1
        return rep
    import torch
    t = torch.tensor([1.0 if rep["pass"] else 0.0, -rep["max_rel_err"]], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    rep = dict(rep)
    rep["pass"] = bool(t[0].item() >= 1.0)
    rep["max_rel_err"] = float(-t[1].item())
    rep["ranks_checked"] = ws
    return rep


# ------------------------------------------------------------------------------------------------
# reference arm: CPU restatement of the reference algorithm on the box's host cores (the
# reference's R path cannot run: no R on this image).  Only this leg, cpu_baseline and the parity
# block execute anything under oracle/.
# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank, ws = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from chromvar_b200 import synthetic
    from oracle import chromvar_oracle as orc
    d, name = make_workload(0, args.cells, args.quick, args.workload)
    C = d["counts"]
    P, N = C.shape
    K = len(d["annoptr"]) - 1
    e = np.asarray(C.sum(axis=1)).ravel() / C.sum()
    bg = synthetic.background_matrix(C, d["bias"], NITER, 0.1, 50, seed=20173)
    workers = os.cpu_count() or 1
    S = max(1, min(K, 2 * workers))                       # two annotations per worker and step
    farm = orc.TaskFarm(C, sets_of(d, range(K)), bg, e, workers)      # pool + row-major counts: set up once
    times = []
    rng = np.random.default_rng(4242)
    for s in range(args.warmup + args.steps):
        ks = rng.choice(K, size=S, replace=False)
        t0 = time.perf_counter()
        farm.run(ks)
        dt = time.perf_counter() - t0
        log("[bench/reference] step %d: %d annotations x %d cells in %.2fs" % (s, S, N, dt))
        if s >= args.warmup:
            times.append(dt)
    farm.close()
    tot = float(np.sum(times))
    value = S * N * len(times) / tot
    sample = "%d random annotations of %d x all %d cells per step, fork pool of %d workers kept across steps" % (S, K, N, workers)
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(name, d, args.gpus, BG_KIND),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample,
                         "note": "CPU restatement of the reference algorithm (scipy fp64, per-annotation "
                                 "sparse products incl. the per-annotation colSums, fork pool = "
                                 "MulticoreParam); not a chromVAR R timing: R is not installed here"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_json(out)


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
def issue_peak_tflops(sm_mhz, operand_bits, n_sms=148):
    """Issue ceiling of the tcgen05 pipe for this kernel's instruction shape, from the probe
    (tools/fp4_probe.cu, profiles/r01j): one M128 N256 instruction per SM every 128.1 cycles, K = 64
    (e2m1, kind::mxf4) or 32 (u8, kind::i8) peaks per instruction."""
    kk = 64 if operand_bits == 4 else 32
    return n_sms * (2.0 * 128 * 256 * kk / 128.1) * sm_mhz * 1e6 / 1e12


def roofline_of(st, c_ms, total_ms_per_step, flops_alg, peaks, peaks_kind, sm_mhz, workload, bytes_alg=0, upd=0,
                shape=None):
    """Roofline of the dominant kernel.  Dense path: the GEMM is bound by the tensor pipe when its
    arithmetic intensity (algorithmic ops per compulsory operand byte of one launch) is above the ridge,
    by HBM otherwise (few cells: every launch streams the whole stacked operand M once -- config 4)."""
    share = c_ms / total_ms_per_step if total_ms_per_step else None
    if st["path"] == 1:
        fp4 = st.get("operand_bits") == 4
        few = bool(st.get("few_cells"))
        launches = 1 if few else max(1, st["n_cell_tiles"] * max(st["planes"], 1))
        achieved = flops_alg / (c_ms * 1e-3) / 1e12
        peak_issue = issue_peak_tflops(sm_mhz or 1965.0, 4 if fp4 else 8)
        mult = 4.0 if fp4 else 2.0
        kname = "k_gather_contract" if few else "k_dense_contract_fp4" if fp4 else (
            "k_dense_contract_pair" if st["cell_tile"] == 256 else "k_dense_contract")
        tensor = {"bound": "tensor", "kernel": kname, "achieved": achieved, "peak": peak_issue, "unit": "TFLOP/s",
                  "frac": achieved / peak_issue,
                  "peak_kind": "probe-measured tcgen05 issue rate of this instruction shape (%s, 128.1 cycles per "
                               "M128 N256 K%d instruction and SM) x 148 SMs x the SM clock sampled in this run (%.0f MHz)"
                               % ("kind::mxf4 e2m1" if fp4 else "kind::i8 u8", 64 if fp4 else 32, sm_mhz or 1965.0),
                  "peak_bf16_derived_burst": mult * peaks["bf16_tflops"],
                  "frac_vs_bf16_derived_burst": achieved / (mult * peaks["bf16_tflops"]),
                  "peak_bf16_derived_kind": peaks_kind + " MEASURED_PEAKS bf16 burst x %d (%s vs bf16 MAC rate); a cuBLAS-bf16-"
                                            "derived figure under-states this pipe for 97 %%-zero operands (no power throttling)"
                                            % (int(mult), "e2m1" if fp4 else "int8"),
                  "peak_nominal": 9000.0 if fp4 else 4500.0, "frac_vs_nominal": achieved / (9000.0 if fp4 else 4500.0)}
        common = {"traffic": ncu_traffic(kname, workload, c_ms / launches),
                  "ms_per_launch": c_ms / launches, "launches_per_step": int(launches),
                  "algorithmic_ops_per_launch": flops_alg / launches, "padded_ops_per_step": st["contract_flops"],
                  "share_of_step": share, "operand_bits": st.get("operand_bits"),
                  "overflow_columns_per_step": st.get("overflow_columns")}
        if few:
            tensor["note"] = ("few-cells contraction: counts rows gathered per iteration into an MN-major operand, both u8 digit "
                              "planes in one pass, no stacked operand; the kernel span includes the finishing kernel")
        if shape is not None and not few:
            P, N, K = shape
            I = NITER + 1
            if fp4:
                rows = -(-K // (256 // I + 208 // I)) * 464
                row_bytes = (-(-P // 256) * 256 + st.get("overflow_columns", 0) / max(1, st["n_cell_tiles"])) / 2.0
            else:
                rows = -(-K // max(1, 256 // I)) * 256
                row_bytes = float(-(-P // 128) * 128)
            m_bytes = rows * row_bytes                               # the stacked operand M, read once per launch
            cells_per_launch = N / max(1, st["n_cell_tiles"])
            c_bytes = -(-int(cells_per_launch) // 256) * 256 * row_bytes
            launch_bytes = m_bytes + c_bytes + 2.0 * 8.0 * K * cells_per_launch / max(st["planes"], 1)
            intensity = (flops_alg / launches) / launch_bytes
            ridge = peak_issue * 1e12 / (peaks["hbm_gbs"] * 1e9)
            common.update({"compulsory_bytes_per_launch": int(launch_bytes), "ops_per_compulsory_byte": intensity,
                           "ridge_ops_per_byte": ridge})
            if intensity < ridge:
                gbs = launch_bytes / (c_ms / launches * 1e-3) / 1e9
                out = {"bound": "hbm", "kernel": kname, "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                       "frac": gbs / peaks["hbm_gbs"],
                       "peak_kind": peaks_kind + " MEASURED_PEAKS copy bandwidth; the launch streams the whole stacked operand "
                                    "M once for a single 256-cell unit: %.0f algorithmic ops per compulsory byte, below the "
                                    "ridge of %.0f" % (intensity, ridge)}
                out.update(common)
                out["tensor_view"] = tensor
                return out
        tensor.update(common)
        return tensor
    achieved = bytes_alg / (c_ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "k_contract", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / peaks["hbm_gbs"], "traffic": None, "peak_kind": peaks_kind, "ms_per_launch": c_ms,
            "algorithmic_bytes_per_launch": int(bytes_alg), "updates_per_launch": int(upd), "share_of_step": share}


def dtype_string(st):
    if st["path"] != 1:
        return "int32 sums, f64 epilogue"
    return "%s, %s epilogue" % ("e2m1 x e2m1 -> f32 holding exact integers (tcgen05 kind::mxf4)" if st.get("operand_bits") == 4
                                else "u8 x u8 -> s32 (tcgen05 kind::i8)",
                                "u64 fixed-point mean / f32 two-sweep variance / f64 tail" if st.get("epilogue") == 32 else "f64")


class Timed:
    """K steps of `step()` on the engine's stream, each bracketed by CUDA events, L2 flushed between."""

    def __init__(self, torch, eng, dev):
        self.torch, self.eng, self.dev = torch, eng, dev
        self.ext = torch.cuda.ExternalStream(eng.stream(), device=dev)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def run(self, step, steps):
        torch = self.torch
        pending = []
        for _ in range(steps):
            with torch.cuda.stream(self.ext):
                self.flush.zero_()                           # L2 flush between timed iterations
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(self.ext)
            step()
            ev1.record(self.ext)
            pending.append((ev0, ev1))
        self.eng.deviations_wait()                           # flag words of the queued calls verified here
        torch.cuda.synchronize(self.dev)
        return [a.elapsed_time(b) for a, b in pending]


def run_workload(torch, dist, eng, dev, rank, ws, d, name, args, steps, warmup, with_e2e, gloo_group=None, parity_anno=None,
                 full_cells_parity=True, bg_from="synthetic"):
    """One workload on this rank's engine: set-up, K device-timed steps, (optionally) the e2e legs,
    parity.  Returns a dict of measurements (every rank; rank 0 assembles the JSON)."""
    from chromvar_b200 import _lib, distributed as cvd, synthetic
    C = d["counts"]
    P, N = C.shape
    K = len(d["annoptr"]) - 1
    nnz = int(C.nnz)
    timed = Timed(torch, eng, dev)
    ext = timed.ext

    def pin(a):
        """Pinned copy of a (memory order kept: the background matrix stays column-major); arrays
        the device generator already left in pinned memory are used as they are."""
        a = np.asarray(a)
        if a.ndim == 1 and a.flags.c_contiguous and a.size and torch.from_numpy(a).is_pinned():
            return a
        if a.ndim == 2 and a.flags.f_contiguous and not a.flags.c_contiguous:
            return torch.from_numpy(np.ascontiguousarray(a.T)).pin_memory().numpy().T
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()

    h_colptr, h_rowidx, h_x = (pin(C.indptr.astype(np.int64)), pin(C.indices.astype(np.int32, copy=False)),
                               pin(C.data.astype(np.float64, copy=False)))
    h_ptr, h_idx = pin(d["annoptr"]), pin(d["annoidx"].astype(np.int32))

    def barrier():
        if ws > 1:
            dist.barrier(group=gloo_group) if gloo_group is not None else dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- setup (untimed): counts -> HBM, global per-peak totals, expectation, background -------
    eng.set_option("check_zero_peaks", 1 if ws == 1 else 0)   # a shard may hold no fragment of a peak; e is global
    if "dense" in d and ws == 1:
        eng.set_counts_dense(d["dense"])                      # bulk: base-matrix input as the reference holds it
    else:
        eng.set_counts_csc(P, N, h_colptr, h_rowidx, h_x)
    cvd.allreduce_peak_totals(eng)                       # NCCL all-reduce of int64[P+1] when ws > 1
    t0 = time.perf_counter()
    e = eng.expectations()
    t_exp = time.perf_counter() - t0
    eng.set_seed(20173)
    eng.background_peaks(d["bias"], NITER, 0.1, 50)      # warm-up (allocations, module load)
    t0 = time.perf_counter()
    bg_gpu = eng.background_peaks(d["bias"], NITER, 0.1, 50)
    t_bg = time.perf_counter() - t0
    bg_ms_dev = eng.stats()["ms_background"]
    other = {"getBackgroundPeaks_ms_device": bg_ms_dev, "getBackgroundPeaks_ms_wall": 1e3 * t_bg,
             "getBackgroundPeaks_peaks": int(P), "sampled_indices_per_s": P * NITER / max(t_bg, 1e-9),
             "getBackgroundPeaks_hbm_frac_of_write": (4.0 * P * NITER + 16.0 * P) / max(bg_ms_dev * 1e-3, 1e-12) / 1e9 /
             peaks_json()[0]["hbm_gbs"], "computeExpectations_ms_wall": 1e3 * t_exp}
    if bg_from == "synthetic" and ws == 1:
        # the matrix both arms are fed (see workload_config): host generative sampler, seeded
        bg = synthetic.background_matrix(C, d["bias"], NITER, 0.1, 50, seed=20173)
        bg_kind = BG_KIND
    else:
        bg = bg_gpu                                       # shards: the GPU sampler on the reduced totals (same seed on every rank)
        bg_kind = "cv_background_peaks (GPU sampler, seed 20173%s)" % (", global totals" if ws > 1 else "")
    h_bg, h_e = pin(bg), pin(e)
    eng.deviations_upload(h_ptr, h_idx, h_bg, h_e, index_base=1)

    # ---- the timed step: every kernel of the path, queued without host synchronisation ------------
    # cv_deviations_compute_async replays the verdicts of the checked warm-up call; with several ranks
    # reduction 2 (variability partials) stays in every step, on the device: NCCL all-gather of the
    # K x 4 partials behind the step's kernels + cv_variability_merge_device.  Same definition at every N.
    eng.deviations_compute(rebuild_counts_layout=True, keep_sums=False)      # the checked call: verdicts, buffers
    mine = gathered = None
    if ws > 1:
        mine = torch.as_tensor(eng.variability_partials_device(), device=dev)    # the engine's own buffer (grow-only)
        gathered = torch.empty(ws * mine.numel(), dtype=mine.dtype, device=dev)

    def step(collective=True):
        eng.deviations_compute_async(rebuild_counts_layout=True)
        if ws > 1 and collective:
            with torch.cuda.stream(ext):
                dist.all_gather_into_tensor(gathered, mine)
            eng.variability_merge_device(gathered.data_ptr(), ws)

    # the clock sampler's set-up (nvmlInit, thread start, first NVML calls: driver locks for milliseconds)
    # happens beside the warm-up steps, not beside the first timed one; its samples start at the timed region
    sampler = ClockSampler(dev.index, args.clock_sampler)
    sampler.start()
    for _ in range(warmup):
        step()
    eng.deviations_wait()
    launches0 = eng.stats()["kernel_launches"]
    barrier()
    sampler.samples = []
    step_ms = timed.run(step, steps)
    barrier()
    launches = eng.stats()["kernel_launches"] - launches0
    st = eng.stats()                                     # timed spans of the last queued step
    total_ms = float(np.sum(step_ms))
    log("[bench] rank %d %s: ms of each timed step: %s" % (rank, name, [round(x, 2) for x in step_ms]))
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if ws > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms_max = float(tmax.item())
    res = {"name": name, "config": workload_config(name, d, ws, bg_kind), "P": P, "N": N, "K": K, "nnz": nnz, "step_ms": step_ms, "total_ms_max": total_ms_max,
           "value": K * N * ws * steps / (total_ms_max * 1e-3), "launches": int(launches), "st": st, "other": other,
           "bg_kind": bg_kind, "steps": steps, "warmup": warmup}
    if ws > 1:
        sd_all = eng.deviations_download(vectors_only=True)["variability"]
        res["variability_merged_on_device"] = bool(sd_all.shape[0] == K and np.isfinite(sd_all).any())

    # ---- e2e through the C ABI: one cv_deviations_csc call per step on HOST buffers ----------------
    out = None
    if with_e2e:
        def outbuf(pinned):
            mk = (lambda a: pin(a.ravel(order="K")).reshape((K, N), order="F")) if pinned else (lambda a: a)
            return dict(dev=mk(np.empty((K, N), np.float64, order="F")), z=mk(np.empty((K, N), np.float64, order="F")),
                        matches=np.empty(K), overlap=np.empty(K), variability=np.empty(K))

        def e2e_leg(bufs, o, n_steps):
            cp, ri, x, ap, ai, b, ee = bufs

            def one():
                # compute_deviations_core as the R shim calls it: ONE ABI call on host buffers (counts CSC,
                # annotations, background, expectation in; deviations / z out).  The library overlaps the
                # chunked H2D of the counts and the chunked D2H of the results with the kernels.
                eng.deviations_csc(P, N, cp, ri, x, ap, ai, b, ee, index_base=1, out=o)
                if ws > 1:
                    with torch.cuda.stream(ext):
                        dist.all_gather_into_tensor(gathered, torch.as_tensor(eng.variability_partials_device(), device=dev))
                    eng.variability_merge_device(gathered.data_ptr(), ws)
                    torch.cuda.synchronize(dev)
            one()
            one()
            barrier()
            import gc
            gc.collect()
            gc.disable()             # a gen-2 collection of the harness costs ~10 ms when it lands inside a step
            t0 = time.perf_counter()
            each, lib = [], []
            for _ in range(n_steps):
                t1 = time.perf_counter()
                one()
                each.append(1e3 * (time.perf_counter() - t1))
                lib.append(eng.stats()["ms_h2d"])          # wall clock inside cv_deviations_csc
                if sampler.mode == "inline":
                    sampler.sample()
            barrier()
            dt = time.perf_counter() - t0
            gc.enable()
            te = torch.tensor([dt], dtype=torch.float64, device=dev)
            if ws > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            return float(te.item()), each, lib

        n_e2e = max(1, min(steps, 10)) if not args.e2e_steps else args.e2e_steps

        def best_of(bufs, o, passes=3):
            """The shared boxes deschedule host threads for 10-100 ms at a time (profiles/r01x); the e2e
            call has the host in its loop, so one stall lands in a step.  Each leg is timed `passes`
            times over n_e2e steps and the best pass is reported (all passes are listed)."""
            runs = [e2e_leg(bufs, o, n_e2e) for _ in range(passes if not args.quick else 1)]
            best = min(runs, key=lambda r: r[0])
            return best + ([round(1e3 * r[0] / n_e2e, 2) for r in runs],)

        out = outbuf(True)
        pinned_bufs = (h_colptr, h_rowidx, h_x, h_ptr, h_idx, h_bg, h_e)
        t_pin, each_pin, lib_pin, passes_pin = best_of(pinned_bufs, out)
        e2e_stats = eng.stats()
        # the same call on PAGEABLE host memory: what R hands over (R's vectors are ordinary heap memory)
        pageable = tuple(np.array(a, copy=True, order="K") for a in pinned_bufs)
        out_pg = outbuf(False)
        t_pg, each_pg, lib_pg, passes_pg = best_of(pageable, out_pg)
        host_in = sum(a.nbytes for a in pinned_bufs)
        res["e2e"] = {"value": K * N * ws * n_e2e / t_pin, "unit": UNIT,
                      "h2d_bytes_per_step": int(e2e_stats.get("h2d_bytes", 0)) or int(host_in),
                      "d2h_bytes_per_step": int(2 * K * N * 8 + 3 * K * 8), "host_input_bytes_per_step": int(host_in),
                      "host_memory": "pinned (cudaHostAlloc'ed by torch)",
                      "ms_per_step": 1e3 * t_pin / n_e2e, "steps": n_e2e, "ms_per_step_median": float(np.median(each_pin)),
                      "timing": "best of %d passes of %d steps (host stalls of the shared box land in single steps); "
                                "ms per step of every pass: %s" % (len(passes_pin), n_e2e, passes_pin),
                      "call": "cv_deviations_csc (chunked H2D / kernels / D2H overlapped)",
                      "gemm_kernel": ("k_dense_contract_fp4" if e2e_stats.get("operand_bits") == 4 else "k_dense_contract_pair")
                      if e2e_stats["cell_tile"] == 256 else ("k_dense_contract" if e2e_stats["path"] == 1 else "k_contract"),
                      "cell_chunks": int(e2e_stats["n_cell_tiles"]), "ms_each_step": [round(t, 2) for t in each_pin],
                      "ms_inside_library": [round(t, 2) for t in lib_pin],
                      "device_ms": {k: round(float(e2e_stats[k]), 3) for k in
                                    ("ms_counts_layout", "ms_anno_prep", "ms_contract", "ms_finalize")},
                      "pageable": {"value": K * N * ws * n_e2e / t_pg, "unit": UNIT, "ms_per_step": 1e3 * t_pg / n_e2e,
                                   "ms_per_step_median": float(np.median(each_pg)), "steps": n_e2e,
                                   "ms_per_step_every_pass": passes_pg,
                                   "host_memory": "pageable numpy arrays (what R's .Call passes), outputs pageable too",
                                   "ms_each_step": [round(t, 2) for t in each_pg],
                                   "results_equal_pinned_leg": bool(np.array_equal(out_pg["z"], out["z"], equal_nan=True))}}
    res["clocks"] = sampler.stop()
    res["bg"], res["e"], res["out"] = bg, e, out
    return res


def check_workload_parity(torch, dist, eng, dev, rank, ws, d, res, args, n_anno, full_cells):
    """Parity of a workload's results (see parity_block); also yields the cpu_baseline figure."""
    C = d["counts"]
    P, N = C.shape
    workers = max(1, (os.cpu_count() or 1) // ws)
    if res["out"] is not None:
        got_dev, got_z = res["out"]["dev"], res["out"]["z"]
    else:
        o = eng.deviations_download()
        got_dev, got_z = o["dev"], o["z"]
    if full_cells:
        lo, hi = 0, N
    else:
        rng = np.random.default_rng(77 + rank)
        lo = int(rng.integers(0, max(1, N - 2048 + 1)))
        hi = min(N, lo + 2048)
    rep, dt, units = parity_block(eng, d, res["bg"], res["e"], got_dev, got_z, rank, workers, n_anno, lo, hi)
    cpu = {"value": units / dt, "unit": UNIT, "cores": workers, "kind": "port",
           "sample": "%d random annotations of %d x %d of %d cells, fork pool of %d workers, %.1fs"
                     % (rep["annotations_checked"], len(d["annoptr"]) - 1, hi - lo, N, workers, dt)}
    return reduce_parity(rep, ws, dist, dev), cpu


def sub_record(res, parity, cpu, peaks, peaks_kind, extra=None):
    """A compact record of one extra workload (configs 3 / 4 / 5-shard, atlas)."""
    st = res["st"]
    P, N, K = res["P"], res["N"], res["K"]
    ms_step = res["total_ms_max"] / res["steps"]
    sm_mhz = (res.get("clocks") or {}).get("sm_mhz")
    roof = roofline_of(st, st["ms_contract"], float(np.mean(res["step_ms"])), 2.0 * (NITER + 1) * K * P * N, peaks, peaks_kind,
                       sm_mhz, res["name"], st["contract_bytes"], st["contract_updates"], shape=(P, N, K))
    rec = {"workload": res["name"], "peaks": P, "cells_per_gpu": N, "annotations": K, "nnz_per_gpu": res["nnz"],
           "value": res["value"], "unit": UNIT, "ms_per_step": ms_step, "steps": res["steps"], "warmup": res["warmup"],
           "ms_each_step": [round(x, 2) for x in res["step_ms"]], "dtype": dtype_string(st),
           "stages_ms": {"counts_layout": st["ms_counts_layout"], "anno_prep": st["ms_anno_prep"],
                         "contract": st["ms_contract"], "finalize": st["ms_finalize"]},
           "digit_planes": st["planes"], "gpu_launches": res["launches"], "roofline": roof, "parity": parity,
           "cpu_baseline": cpu, "background": res["bg_kind"], "other_paths": res["other"], "clocks": res.get("clocks")}
    if extra:
        rec.update(extra)
    return rec


def run_ours(args):
    import torch
    import torch.distributed as dist
    from chromvar_b200 import _lib, distributed as cvd

    rank, ws, local = cvd.init()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    gloo = dist.new_group(backend="gloo") if ws > 1 else None       # host-side barriers that keep the GPUs free
    eng = _lib.Handle(local, seed=20173)
    eng.set_option("path", args.path)
    eng.set_option("dense_cta_pair", args.pair)
    eng.set_option("dense_fp4", args.fp4)
    eng.set_option("dense_lockstep", args.lockstep)
    eng.set_option("host_narrow", args.host_narrow)
    eng.set_option("tail_chunk", args.tail_chunk)
    if args.chunk:
        eng.set_option("cell_chunk", args.chunk)
    if args.band:
        eng.set_option("dense_band", args.band)
    if args.debug_no_tma:
        eng.set_option("debug_no_tma", args.debug_no_tma)
    peaks, peaks_kind = peaks_json()
    d, name = make_workload(rank, args.cells, args.quick, args.workload, gen=args.gen, device="cuda:%d" % local)
    P, N = d["counts"].shape
    K = len(d["annoptr"]) - 1

    main = run_workload(torch, dist, eng, dev, rank, ws, d, name, args, args.steps, args.warmup, with_e2e=True,
                        gloo_group=gloo)
    st = main["st"]

    # ---- the R call path on all GPUs of this node from ONE process (cv_multi), rank 0 only -----------
    single = None
    if ws > 1 and not args.no_single_process:
        dist.barrier(group=gloo)
        if rank == 0:
            try:
                single = single_process_leg(torch, d, main, ws, args)
            except Exception as ex:                      # reported, never fatal for the main line
                single = {"error": "%s: %s" % (type(ex).__name__, ex)}
            log("[bench] single-process leg: %s" % {k: v for k, v in (single or {}).items() if k != "ms_each_step"})
        dist.barrier(group=gloo)

    # ---- parity + CPU baseline of the main workload (changes the resident counts: after the timing) ---
    parity = cpu = None
    if not args.no_cpu_baseline:
        n_anno = 2 * max(1, (os.cpu_count() or 1) // ws) if main["nnz"] < 1e8 else max(4, (os.cpu_count() or 1) // ws)
        parity, cpu = check_workload_parity(torch, dist, eng, dev, rank, ws, d, main, args, n_anno,
                                            full_cells=main["nnz"] < 1e8)
        log("[bench] rank %d parity of the timed e2e results vs the CPU restatement: %s" % (rank, parity))
    del main["out"]

    # ---- other BASELINE configurations, driver-visible (N = 1: configs 3, 4, 5-shard; N > 1: the atlas) --
    subs = {}
    if not args.quick and args.workload is None and not args.no_sub_workloads:
        del d
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        names = ["cfg3_kmers", "cfg4_bulk", "cfg5_atlas_shard"] if ws == 1 else ["cfg5_atlas_shard"]
        for wname in names:
            try:
                t_sub = time.time()
                dd, _ = make_workload(rank, 0, False, wname, gen="auto", device="cuda:%d" % local)
                r = run_workload(torch, dist, eng, dev, rank, ws, dd, wname, args, args.sub_steps, 1, with_e2e=False,
                                 gloo_group=gloo, bg_from="gpu")
                extra = {}
                if ws > 1:
                    # the same shard on rank 0 alone (the other GPUs idle): atlas speed-up on THIS box
                    dist.barrier(group=gloo)
                    if rank == 0:
                        tm = Timed(torch, eng, dev)
                        solo = tm.run(lambda: eng.deviations_compute_async(rebuild_counts_layout=True), args.sub_steps)
                        solo_value = r["K"] * r["N"] * args.sub_steps / (float(np.sum(solo)) * 1e-3)
                        extra = {"one_gpu_shard_value_same_box": solo_value, "one_gpu_shard_ms_each_step": [round(x, 2) for x in solo],
                                 "speedup_vs_one_gpu_shard": r["value"] / solo_value, "cells_total": r["N"] * ws,
                                 "scaling": "weak: one config-5 shard (125k cells) per GPU, %d GPUs = %d cells" % (ws, r["N"] * ws)}
                    dist.barrier(group=gloo)
                par, cpu_s = (None, None)
                if not args.no_cpu_baseline:
                    par, cpu_s = check_workload_parity(torch, dist, eng, dev, rank, ws, dd, r, args,
                                                       n_anno=max(4, min(8, (os.cpu_count() or 1) // ws)), full_cells=wname == "cfg4_bulk")
                key = {"cfg3_kmers": "cfg3", "cfg4_bulk": "cfg4", "cfg5_atlas_shard": "cfg5_shard" if ws == 1 else "atlas"}[wname]
                if rank == 0:
                    subs[key] = sub_record(r, par, cpu_s, peaks, peaks_kind, extra)
                    subs[key]["wall_s_incl_generation"] = round(time.time() - t_sub, 1)
                    log("[bench] %s: %.3g z/s, %.1f ms/step, parity %s" % (key, r["value"], r["total_ms_max"] / r["steps"], par))
                del dd, r
                gc.collect()
                torch.cuda.empty_cache()
            except Exception as ex:
                if rank == 0:
                    subs[wname] = {"error": "%s: %s" % (type(ex).__name__, ex)}
                log("[bench] rank %d: sub-workload %s failed: %s" % (rank, wname, ex))

    if rank != 0:
        return
    sm_mhz = (main["clocks"] or {}).get("sm_mhz")
    c_ms = st["ms_contract"]
    roof = roofline_of(st, c_ms, float(np.mean(main["step_ms"])), 2.0 * (NITER + 1) * K * P * N, peaks, peaks_kind, sm_mhz,
                       name, st["contract_bytes"], st["contract_updates"], shape=(P, N, K))
    line = {
        "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": ws, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": main["total_ms_max"] / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": dtype_string(st), "data": "synthetic",
        "config": main["config"],
        "engine": {"cell_tile": int(st["cell_tile"]),
                   "step_call": "cv_deviations_compute_async: K steps queued back to back, settled after the loop"
                                + ("; per step an NCCL all-gather of the K x 4 variability partials + cv_variability_merge_device, "
                                   "stream-ordered" if ws > 1 else ""),
                   "path": ("dense_tcgen05_cta_pair" if st["cell_tile"] == 256 else "dense_tcgen05") if st["path"] == 1
                   else "row_gather_spmm"},
        "e2e": main["e2e"],
        "gpu_launches": main["launches"],
        "clocks": main["clocks"],
        "roofline": roof,
        "cpu_baseline": cpu if ws == 1 else None,
        "parity": parity,
        "stages_ms": {"counts_layout": st["ms_counts_layout"], "anno_prep": st["ms_anno_prep"],
                      "contract": st["ms_contract"], "finalize": st["ms_finalize"]},
        "ms_each_step": [round(x, 2) for x in main["step_ms"]],
        "other_paths": main["other"],
        "workloads": subs or None,
        "single_process": single,
    }
    if ws > 1 and "variability_merged_on_device" in main:
        line["variability_merged_on_device"] = main["variability_merged_on_device"]
    emit_json(line)


def single_process_leg(torch, d, main, ws, args):
    """compute_deviations_core through cv_multi: ONE process (this one) drives all `ws` GPUs of the
    node -- the call the R shim makes (cvb_deviations_csc).  The counts are `ws` column blocks (this
    rank's matrix repeated: the same per-GPU work as the multi-process line), host memory pageable."""
    from chromvar_b200 import _lib
    C = d["counts"]
    P, N = C.shape
    K = len(d["annoptr"]) - 1
    m = _lib.MultiHandle(list(range(ws)), seed=20173)
    try:
        for key, v in (("path", args.path), ("dense_cta_pair", args.pair), ("dense_fp4", args.fp4), ("tail_chunk", args.tail_chunk)):
            m.set_option(key, v)
        blocks = m._blocks([C] * ws)
        out = dict(dev=np.empty((K, N * ws), np.float64, order="F"), z=np.empty((K, N * ws), np.float64, order="F"),
                   matches=np.empty(K), overlap=np.empty(K), variability=np.empty(K))
        ptr, idx = d["annoptr"], d["annoidx"].astype(np.int32)
        bg, e = main["bg"], main["e"]

        def one():
            m.deviations_csc(None, ptr, idx, bg, e, index_base=1, out=out, blocks=blocks)
        one()
        one()
        n = max(1, min(args.steps, 10))
        each = []
        t0 = time.perf_counter()
        for _ in range(n):
            t1 = time.perf_counter()
            one()
            each.append(1e3 * (time.perf_counter() - t1))
        dt = time.perf_counter() - t0
        same = bool(main["out"] is not None and np.array_equal(out["z"][:, :N], main["out"]["z"], equal_nan=True))
        reps = bool(all(np.array_equal(out["z"][:, b * N:(b + 1) * N], out["z"][:, :N], equal_nan=True) for b in range(1, ws)))
        return {"call": "cv_multi_deviations_csc: one process, one host thread per GPU inside the library, counts as %d column "
                        "blocks, pageable host memory, reductions over peer memory" % ws,
                "n_gpus": ws, "peer_access": m.peer_access(), "value": K * N * ws * n / dt, "unit": UNIT,
                "ms_per_step": 1e3 * dt / n, "ms_per_step_median": float(np.median(each)), "steps": n,
                "ms_each_step": [round(t, 2) for t in each], "shards": [int(x) for x in m.shards()],
                "z_equals_multi_process_result_bitwise": same, "z_equal_across_shards_bitwise": reps}
    finally:
        m.close()


_REAL_STDOUT = None


def quiet_stdout():
    """stdout carries exactly one JSON line: anything a library prints there meanwhile (NCCL's
    version banner, for one) goes to stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_json(obj):
    sys.stdout.flush()
    data = (json.dumps(obj) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=0, help="override cells per GPU (default: config 2's 10000)")
    ap.add_argument("--workload", default=None, help="synthetic configuration (default cfg2_scatac_motifs; "
                    "cfg3_kmers, cfg4_bulk, cfg5_atlas_shard are parity / scale cases, not the bench line)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="number of timed e2e steps (default min(steps, 10))")
    ap.add_argument("--gen", default="auto", choices=["auto", "host", "device"],
                    help="synthetic counts on the host (numpy) or on the GPU (same model, torch); auto: GPU for >= 5e8 nonzeros")
    ap.add_argument("--host-narrow", type=int, default=-1, help="e2e call: -1 (library default) host threads stage / narrow PAGEABLE caller memory only, 0 never, 1 always, n > 1: with n threads")
    ap.add_argument("--quick", action="store_true", help="small workload (development only; not a bench value)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--path", type=int, default=0, help="0 auto (cost model), 1 row-gather SpMM, 2 dense tcgen05")
    ap.add_argument("--lockstep", type=int, default=1, help="e2m1 GEMM: producers of all CTAs start each sub-tile together")
    ap.add_argument("--fp4", type=int, default=-1, help="dense path: e2m1 operands when representable (-1 / 1) or never (0)")
    ap.add_argument("--chunk", type=int, default=0, help="dense path: cap on cells per GEMM launch (experiment)")
    ap.add_argument("--band", type=int, default=0, help="dense path: groups per raster band (experiment)")
    ap.add_argument("--debug-no-tma", type=int, default=0, help="experiment: GEMM without operand loads (1) / MMA issue patterns (2-5); NOT a bench value")
    ap.add_argument("--pair", type=int, default=-1, help="dense GEMM kernel: 1 CTA pair (cta_group::2), 0 single CTA, -1 library default (by context)")
    ap.add_argument("--clock-sampler", default="thread", choices=["thread", "inline", "off"],
                    help="SM clock / throttle reasons through NVML: polled by a thread every 100 ms, or by the main thread between steps")
    ap.add_argument("--tail-chunk", type=int, default=1, help="e2e call: 1 (default) ends with a short fourth cell chunk, 0: three chunks")
    ap.add_argument("--sub-steps", type=int, default=3, help="timed steps of each extra workload (configs 3 / 4 / 5-shard, atlas)")
    ap.add_argument("--no-sub-workloads", action="store_true", help="only the main workload (config 2)")
    ap.add_argument("--no-single-process", action="store_true", help="N > 1: skip the one-process cv_multi leg")
    ap.add_argument("--pin-cores", type=int, default=0,
                    help="multi-GPU experiment: 1 keeps each rank on its own block of host cores (cores / ranks, by "
                         "local rank); measured on 8x B200 / 32 cores: no gain for the e2e call, device-timed step slower")
    args = ap.parse_args()
    ws_env, local_env = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "ours" and ws_env > 1 and args.pin_cores != 0 and hasattr(os, "sched_setaffinity"):
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // ws_env)
            os.sched_setaffinity(0, cores[local_env * per:(local_env + 1) * per] or cores)
        except OSError:
            pass
    if args.warmup < 3 and not args.quick and args.impl == "ours":
        log("[bench] warning: fewer than 3 warm-up steps")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
        try:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()
        except Exception:
            pass


if __name__ == "__main__":
    main()
