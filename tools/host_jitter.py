"""Host scheduling jitter probe: largest gaps seen by a busy loop (no CUDA, no threads)."""
import time, os
t_end = time.perf_counter() + 3.0
last = time.perf_counter(); gaps = []
while last < t_end:
    now = time.perf_counter()
    if now - last > 0.002: gaps.append(round(1e3 * (now - last), 1))
    last = now
print("cpus", os.cpu_count(), "gaps > 2 ms in 3 s:", len(gaps), sorted(gaps)[-8:])
try:
    print(open("/sys/fs/cgroup/cpu.max").read().strip(), "|", [l for l in open("/sys/fs/cgroup/cpu.stat").read().split("\n") if "thrott" in l])
except Exception as e:
    print("cgroup:", e)
