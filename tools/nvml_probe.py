"""Which NVML query stalls CUDA submissions?  Runs the bench step in a loop while a thread polls one
NVML call type every 100 ms; prints step-time outliers and the latency of the NVML calls themselves."""
import os, sys, time, threading, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pynvml
from chromvar_b200 import _lib, synthetic

pynvml.nvmlInit()
dev = pynvml.nvmlDeviceGetHandleByIndex(0)
d = synthetic.make_config("cfg2_scatac_motifs")
C = d["counts"]; P, N = C.shape
eng = _lib.Handle(0, seed=1)
eng.set_counts_csc(P, N, C.indptr.astype(np.int64), C.indices.astype(np.int32), C.data)
e = eng.expectations()
bg = eng.background_peaks(d["bias"], 50, 0.1, 50)
eng.deviations_upload(d["annoptr"], d["annoidx"].astype(np.int32), bg, e, index_base=1)
for _ in range(3):
    eng.deviations_compute(rebuild_counts_layout=True, keep_sums=False)

calls = {
    "none": [],
    "clock_info": [lambda: pynvml.nvmlDeviceGetClockInfo(dev, pynvml.NVML_CLOCK_SM)],
    "event_reasons": [lambda: pynvml.nvmlDeviceGetCurrentClocksEventReasons(dev)],
    "both": [lambda: pynvml.nvmlDeviceGetClockInfo(dev, pynvml.NVML_CLOCK_SM), lambda: pynvml.nvmlDeviceGetCurrentClocksEventReasons(dev)],
    "none_again": [],
}
period = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
for name, fns in calls.items():
    stop = threading.Event(); lat = []
    def poll():
        while not stop.is_set():
            for f in fns:
                t = time.perf_counter(); f(); lat.append(1e3 * (time.perf_counter() - t))
            stop.wait(period)
    th = threading.Thread(target=poll, daemon=True)
    if fns: th.start()
    steps = []
    for _ in range(60):
        t = time.perf_counter()
        eng.deviations_compute(rebuild_counts_layout=True, keep_sums=False)
        steps.append(1e3 * (time.perf_counter() - t))
    stop.set()
    if fns: th.join()
    out = [round(x, 1) for x in steps if x > 18.5]
    print("%-14s step mean %.2f med %.2f  outliers %d (sum +%.0f ms) %s | nvml calls %d, latency med %.2f max %.1f ms" %
          (name, statistics.mean(steps), statistics.median(steps), len(out), sum(out) - 16.6 * len(out), out[:8],
           len(lat), statistics.median(lat) if lat else 0, max(lat) if lat else 0), flush=True)
