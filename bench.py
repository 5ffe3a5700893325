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
          result matrices.
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


def ncu_traffic(kernel, workload):
    """dram bytes (read + write) per launch of `kernel` on `workload` from the committed ncu --set
    full capture (profiles/traffic.json, written by hand from the .ncu-rep), or None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t.get(workload, {}).get(kernel)
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
        big = cfg["density"] < 1.0 and cfg["P"] * cfg["N"] * cfg["density"] >= 5e8
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


# ------------------------------------------------------------------------------------------------
# reference arm: CPU restatement of the reference algorithm (the reference's R path cannot run:
# no R on this image).  Only this leg and cpu_baseline execute anything under oracle/.
# ------------------------------------------------------------------------------------------------
def cpu_reference_step(d, bg, e, anno_slice, workers, want_results=False):
    from oracle import chromvar_oracle as orc
    ptr, idx = d["annoptr"], d["annoidx"]
    sets = [idx[ptr[k]:ptr[k + 1]].astype(np.int64) for k in anno_slice]
    t0 = time.perf_counter()
    res = orc.deviations_task_farm(d["counts"], sets, bg, e, workers)
    dt = time.perf_counter() - t0
    return (dt, res) if want_results else dt


def parity_report(got_dev, got_z, ref_dev, ref_z):
    """GPU results of the timed e2e step against the CPU restatement on the sampled annotations:
    abs(d) <= 1e-5 * max(abs(ref), 1e-3) and identical NA masks."""
    out = {"annotations_checked": int(ref_z.shape[0]), "cells": int(ref_z.shape[1]), "tolerance": 1e-5}
    worst = 0.0
    ok = True
    for g, r in ((got_dev, ref_dev), (got_z, ref_z)):
        ng, nr = np.isnan(g), np.isnan(r)
        ok &= bool(np.array_equal(ng, nr))
        m = ~(nr | ng | np.isinf(r))
        if m.any():
            worst = max(worst, float((np.abs(g[m] - r[m]) / np.maximum(np.abs(r[m]), 1e-3)).max()))
    out["max_rel_err"] = worst
    out["na_masks_equal"] = ok
    out["pass"] = bool(ok and worst <= 1e-5)
    return out


def host_background(d, rng):
    """A background matrix for the reference arm when no GPU sampler is available: peaks with a
    similar fragment total (what the sampler's bins encode), cheap and deterministic."""
    fpp = np.asarray(d["counts"].sum(axis=1)).ravel()
    order = np.argsort(fpp, kind="stable")
    rank_of = np.empty_like(order)
    rank_of[order] = np.arange(order.size)
    P = fpp.size
    jitter = rng.integers(-200, 201, size=(P, NITER))
    pos = np.clip(rank_of[:, None] + jitter, 0, P - 1)
    return np.asfortranarray((order[pos] + 1).astype(np.int32))


def run_reference(args):
    rank, ws = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    d, name = make_workload(0, args.cells, args.quick, args.workload)
    C = d["counts"]
    P, N = C.shape
    K = len(d["annoptr"]) - 1
    e = np.asarray(C.sum(axis=1)).ravel() / C.sum()
    bg = host_background(d, np.random.default_rng(20173))
    workers = os.cpu_count() or 1
    S = max(1, min(K, workers))
    times = []
    for s in range(args.warmup + args.steps):
        sl = [(s * S + j) % K for j in range(S)]
        dt = cpu_reference_step(d, bg, e, sl, workers)
        log("[bench/reference] step %d: %d annotations x %d cells in %.2fs" % (s, S, N, dt))
        if s >= args.warmup:
            times.append(dt)
    tot = float(np.sum(times))
    value = S * N * len(times) / tot
    sample = "%d of %d annotations x all %d cells per step, fork pool of %d workers" % (S, K, N, workers)
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": name, "peaks": int(P), "cells_per_gpu": int(N), "annotations": int(K),
                   "niterations": NITER, "nnz": int(C.nnz)},
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
def run_ours(args):
    import torch
    import torch.distributed as dist
    from chromvar_b200 import _lib, distributed as cvd

    rank, ws, local = cvd.init()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
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
    d, name = make_workload(rank, args.cells, args.quick, args.workload, gen=args.gen, device="cuda:%d" % local)
    C = d["counts"]
    P, N = C.shape
    K = len(d["annoptr"]) - 1
    nnz = int(C.nnz)

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

    # ---- setup (untimed): counts -> HBM, global per-peak totals, expectation, background -------
    eng.set_counts_csc(P, N, h_colptr, h_rowidx, h_x)
    cvd.allreduce_peak_totals(eng)                       # NCCL all-reduce of int64[P+1] when ws > 1
    t0 = time.perf_counter()
    e = eng.expectations()
    t_exp = time.perf_counter() - t0
    eng.background_peaks(d["bias"], NITER, 0.1, 50)      # warm-up (allocations, module load)
    t0 = time.perf_counter()
    bg = eng.background_peaks(d["bias"], NITER, 0.1, 50)
    t_bg = time.perf_counter() - t0
    bg_ms_dev = eng.stats()["ms_background"]
    # the sampler of makePermutedSets (get_background_peaks_alternative, R/test_sets.R:291-330): exact
    # 10 nearest neighbours of every peak in the whitened (fragments, bias) plane + the draws
    eng.background_peaks_knn(d["bias"], NITER, 10)
    t0 = time.perf_counter()
    eng.background_peaks_knn(d["bias"], NITER, 10)
    t_knn = time.perf_counter() - t0
    knn_ms_dev = eng.stats()["ms_background"]
    h_bg, h_e = pin(bg), pin(e)
    eng.deviations_upload(h_ptr, h_idx, h_bg, h_e, index_base=1)

    ext = torch.cuda.ExternalStream(eng.stream(), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # One rank: the K timed steps are enqueued back to back through cv_deviations_compute_async (every
    # kernel of the path runs each step; the host-side verdicts are those of the checked warm-up calls),
    # so a descheduled host thread -- these shared boxes stall the host for 10-100 ms at a time -- no
    # longer shows up as GPU idle time inside a step.  Several ranks: the variability reduction needs
    # each step's partials on the host, so the checked call and its synchronisation stay.
    use_async = args.async_steps == 1 if ws == 1 else args.async_steps == 2
    gathered = mine = None

    def step():
        if use_async:
            eng.deviations_compute_async(rebuild_counts_layout=True)
            if ws > 1:
                # reduction 2 stays in every step, on the device: NCCL all-gather of the K x 4 partials
                # behind the step's kernels (stream-ordered, no host wait); Chan merge after the loop
                with torch.cuda.stream(ext):
                    dist.all_gather_into_tensor(gathered, mine)
            return
        eng.deviations_compute(rebuild_counts_layout=True, keep_sums=False)
        if ws > 1:
            cvd.gather_variability(eng)                  # reduction 2 (K x 4 doubles)

    if use_async and ws > 1:
        eng.deviations_compute(rebuild_counts_layout=True, keep_sums=False)      # the checked call: verdicts, buffers
        mine = torch.as_tensor(eng.variability_partials_device(), device=dev)    # the engine's own buffer (grow-only)
        gathered = torch.empty(ws * mine.numel(), dtype=mine.dtype, device=dev)
    for _ in range(args.warmup):
        step()
    if use_async:
        eng.deviations_wait()
    launches0 = eng.stats()["kernel_launches"]
    sampler = ClockSampler(local, args.clock_sampler)
    sampler.start()
    barrier()
    step_ms, contract_ms, stage_acc = [], [], np.zeros(4)
    upd = bytes_alg = path = 0
    pending = []
    for _ in range(args.steps):
        with torch.cuda.stream(ext):
            flush.zero_()                                # L2 flush between timed iterations
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(ext)
        step()
        ev1.record(ext)
        if use_async:
            pending.append((ev0, ev1))                   # settled after the loop
            continue
        ev1.synchronize()
        step_ms.append(ev0.elapsed_time(ev1))
        if sampler.mode == "inline":
            sampler.sample()
        st = eng.stats()
        contract_ms.append(st["ms_contract"])
        stage_acc += [st["ms_counts_layout"], st["ms_anno_prep"], st["ms_contract"], st["ms_finalize"]]
        upd, bytes_alg, path = st["contract_updates"], st["contract_bytes"], st["path"]
        tile_cells, gemm_launches, padded_flops = st["cell_tile"], st["n_cell_tiles"] * max(st["planes"], 1), st["contract_flops"]
    if use_async:
        eng.deviations_wait()                            # flag words of the queued calls verified here
        torch.cuda.synchronize(dev)
        if ws > 1:
            sd_all = _lib.merge_variability(gathered.view(ws, -1, 4).cpu().numpy(), na_rm=True)
            assert sd_all.shape[0] == K and np.isfinite(sd_all).any()
        step_ms = [a.elapsed_time(b) for a, b in pending]
        st = eng.stats()                                 # timed spans of the last queued step
        contract_ms = [st["ms_contract"]]
        stage_acc += np.array([st["ms_counts_layout"], st["ms_anno_prep"], st["ms_contract"], st["ms_finalize"]]) * args.steps
        upd, bytes_alg, path = st["contract_updates"], st["contract_bytes"], st["path"]
        tile_cells, gemm_launches, padded_flops = st["cell_tile"], st["n_cell_tiles"] * max(st["planes"], 1), st["contract_flops"]
    barrier()
    launches = eng.stats()["kernel_launches"] - launches0
    total_ms = float(np.sum(step_ms))
    log("[bench] rank %d: ms of each timed step: %s" % (rank, [round(x, 2) for x in step_ms]))
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if ws > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms_max = float(tmax.item())
    value = K * N * ws * args.steps / (total_ms_max * 1e-3)

    # ---- e2e through the C ABI with pinned host buffers ---------------------------------------------
    out = dict(dev=pin(np.empty((K, N), np.float64, order="F").ravel(order="K")).reshape((K, N), order="F"),
               z=pin(np.empty((K, N), np.float64, order="F").ravel(order="K")).reshape((K, N), order="F"),
               matches=np.empty(K), overlap=np.empty(K), variability=np.empty(K))

    def e2e_step():
        # compute_deviations_core as the R shim calls it: ONE ABI call on host buffers (counts CSC,
        # annotations, background, expectation in; deviations / z out).  The library overlaps the
        # chunked H2D of the counts and the chunked D2H of the results with the kernels.
        eng.deviations_csc(P, N, h_colptr, h_rowidx, h_x, h_ptr, h_idx, h_bg, h_e, index_base=1, out=out)
        if ws > 1:
            cvd.gather_variability(eng)

    if ws > 1:
        eng.set_option("check_zero_peaks", 0)     # a shard may hold no fragment of a peak; e is global

    e2e_step()
    e2e_step()
    barrier()
    import gc
    gc.collect()
    gc.disable()                 # a gen-2 collection of the harness costs ~10 ms when it lands inside a step
    t0 = time.perf_counter()
    n_e2e = max(1, min(args.steps, 10)) if not args.e2e_steps else args.e2e_steps
    e2e_each, e2e_lib = [], []
    for _ in range(n_e2e):
        t1 = time.perf_counter()
        e2e_step()
        e2e_each.append(1e3 * (time.perf_counter() - t1))
        e2e_lib.append(eng.stats()["ms_h2d"])          # wall clock inside cv_deviations_csc
        if sampler.mode == "inline":
            sampler.sample()
    barrier()
    e2e_s = time.perf_counter() - t0
    gc.enable()
    clocks = sampler.stop()                              # sampled across both timed regions
    e2e_stats = eng.stats()
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if ws > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = K * N * ws * n_e2e / float(te.item())
    host_in = h_colptr.nbytes + h_rowidx.nbytes + h_x.nbytes + h_ptr.nbytes + h_idx.nbytes + h_bg.nbytes + h_e.nbytes
    # bytes that crossed PCIe: the library narrows the fp64 values to bytes on host threads before the
    # copy (cv_hostnarrow.cpp), so fewer bytes travel than the host buffers hold
    h2d = int(e2e_stats.get("h2d_bytes", 0)) or host_in
    d2h = 2 * K * N * 8 + 3 * K * 8

    if rank != 0:
        return
    # ---- roofline of the dominant kernel -----------------------------------------------------------
    peaks, peaks_kind = peaks_json()
    c_ms = float(np.mean(contract_ms))
    share = c_ms / (total_ms / args.steps)
    if path == 1:
        # dense tcgen05 path: algorithmic ops = 2 (B+1) K P N, int8 peak = 2 x the
        # measured bf16 figure (sustained: the kernel runs inside a multi-step loop)
        flops = 2.0 * (NITER + 1) * K * P * N
        fp4 = st.get("operand_bits") == 4
        # e2m1 operands (kind::mxf4) retire twice the MACs per cycle of u8: peak = 4 x bf16
        mult = 4.0 if fp4 else 2.0
        peak = mult * peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
        achieved = flops / (c_ms * 1e-3) / 1e12
        kname = "k_dense_contract_fp4" if fp4 else ("k_dense_contract_pair" if tile_cells == 256 else "k_dense_contract")
        roof = {"bound": "tensor", "kernel": kname, "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak, "traffic": ncu_traffic(kname, name),
                "peak_kind": peaks_kind + (" (4 x bf16 sustained = e2m1 dense)" if fp4 else " (2 x bf16 sustained = int8 dense)"),
                "peak_burst": mult * peaks["bf16_tflops"], "peak_nominal": 9000.0 if fp4 else 4500.0,
                "ms_per_launch": c_ms / max(gemm_launches, 1), "launches_per_step": int(gemm_launches),
                "algorithmic_ops_per_launch": flops / max(gemm_launches, 1),
                "padded_ops_per_step": padded_flops, "share_of_step": share,
                "operand_bits": st.get("operand_bits"), "overflow_columns_per_step": st.get("overflow_columns")}
    else:
        achieved = bytes_alg / (c_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "k_contract", "achieved": achieved, "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": None,
                "peak_kind": peaks_kind, "ms_per_launch": c_ms, "algorithmic_bytes_per_launch": int(bytes_alg),
                "updates_per_launch": int(upd), "share_of_step": share}

    # ---- CPU baseline beside it (N=1 only): bounded sample of the same workload ---------------
    cpu = parity = None
    if ws == 1 and not args.no_cpu_baseline:
        workers = os.cpu_count() or 1
        S = max(1, min(K, (2 if nnz < 1e8 else 1) * workers))
        dt, (ref_z, ref_dev) = cpu_reference_step(d, bg, e, list(range(S)), workers, want_results=True)
        cpu = {"value": S * N / dt, "unit": UNIT, "cores": workers, "kind": "port",
               "sample": "%d of %d annotations x all %d cells, fork pool of %d workers, %.1fs" % (S, K, N, workers, dt)}
        parity = parity_report(out["dev"][:S], out["z"][:S], ref_dev, ref_z)
        log("[bench] parity of the timed e2e results vs the CPU restatement: %s" % parity)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ws, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": ("%s, %s epilogue" % ("e2m1 x e2m1 -> f32 holding exact integers (tcgen05 kind::mxf4)" if st.get("operand_bits") == 4
                                       else "u8 x u8 -> s32 (tcgen05 kind::i8)",
                                       "u64 fixed-point / f32 / f64" if st.get("epilogue") == 32 else "f64"))
                 if path == 1 else "int32 sums, f64 epilogue",
        "data": "synthetic",
        "config": {"workload": name, "peaks": int(P), "cells_per_gpu": int(N), "annotations": int(K),
                   "niterations": NITER, "nnz_per_gpu": nnz, "nnz_annotations": int(d["annoptr"][-1]),
                   "cells_total": int(N) * ws, "generator": d["cfg"].get("generator", "numpy (make_counts)"),
                   "l2": "256 MB flush write between timed steps", "cell_tile": int(tile_cells),
                   "step_call": ("cv_deviations_compute_async: K steps queued back to back, settled after the loop"
                                 if use_async else "cv_deviations_compute: one checked call per step"),
                   "path": ("dense_tcgen05_cta_pair" if tile_cells == 256 else "dense_tcgen05") if path == 1
                   else "row_gather_spmm"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "host_input_bytes_per_step": int(host_in),
                "ms_per_step": 1e3 * float(te.item()) / n_e2e, "steps": n_e2e,
                "ms_per_step_median": float(np.median(e2e_each)),
                "call": "cv_deviations_csc (chunked H2D / kernels / D2H overlapped)",
                "gemm_kernel": ("k_dense_contract_fp4" if e2e_stats.get("operand_bits") == 4 else "k_dense_contract_pair") if e2e_stats["cell_tile"] == 256 else
                ("k_dense_contract" if e2e_stats["path"] == 1 else "k_contract"),
                "cell_chunks": int(e2e_stats["n_cell_tiles"]), "ms_each_step": [round(t, 2) for t in e2e_each],
                "ms_inside_library": [round(t, 2) for t in e2e_lib],
                "device_ms": {k: round(float(e2e_stats[k]), 3) for k in
                              ("ms_counts_layout", "ms_anno_prep", "ms_contract", "ms_finalize")}},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "cpu_baseline": cpu,
        "parity": parity,
        "stages_ms": dict(zip(["counts_layout", "anno_prep", "contract", "finalize"],
                              (stage_acc / args.steps).round(4).tolist())),
        "other_paths": {"getBackgroundPeaks_ms_device": bg_ms_dev, "getBackgroundPeaks_ms_wall": 1e3 * t_bg,
                        "sampled_indices_per_s": P * NITER / max(t_bg, 1e-9),
                        "computeExpectations_ms_wall": 1e3 * t_exp,
                        "backgroundPeaksAlternative_window10_ms_device": knn_ms_dev,
                        "backgroundPeaksAlternative_window10_ms_wall": 1e3 * t_knn},
    }
    emit_json(line)


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
    ap.add_argument("--host-narrow", type=int, default=0, help="e2e call: 0 sends the fp64 values as they are (default), 1 narrows them to bytes on host threads, n > 1: with n threads")
    ap.add_argument("--quick", action="store_true", help="small workload (development only; not a bench value)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--path", type=int, default=0, help="0 auto (cost model), 1 row-gather SpMM, 2 dense tcgen05")
    ap.add_argument("--lockstep", type=int, default=1, help="e2m1 GEMM: producers of all CTAs start each sub-tile together")
    ap.add_argument("--fp4", type=int, default=-1, help="dense path: e2m1 operands when representable (-1 / 1) or never (0)")
    ap.add_argument("--chunk", type=int, default=0, help="dense path: cap on cells per GEMM launch (experiment)")
    ap.add_argument("--band", type=int, default=0, help="dense path: groups per raster band (experiment)")
    ap.add_argument("--debug-no-tma", type=int, default=0, help="experiment: GEMM without operand loads (1) / MMA issue patterns (2-5); NOT a bench value")
    ap.add_argument("--pair", type=int, default=-1, help="dense GEMM kernel: 1 CTA pair (cta_group::2), 0 single CTA, -1 library default (by context)")
    ap.add_argument("--async-steps", type=int, default=1,
                    help="1 (default): one rank enqueues the timed steps back to back (cv_deviations_compute_async), several ranks run one checked call + host-side variability merge per step; 2: several ranks queue their steps too (all-gather on the device); 0: one checked call per step")
    ap.add_argument("--clock-sampler", default="thread", choices=["thread", "inline", "off"],
                    help="SM clock / throttle reasons through NVML: polled by a thread every 100 ms, or by the main thread between steps")
    ap.add_argument("--tail-chunk", type=int, default=1, help="e2e call: 1 (default) ends with a short fourth cell chunk, 0: three chunks")
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
