"""Per-phase wall times of bench.py's e2e step (why is e2e slower than the kernel?)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robotics-path-planning_b200"))
import numpy as np
import torch
import rrtk
from rrtk import workloads as W
cfg = W.C2
Q, iters, n_obs = 4096, 2000, 256
qids = list(range(Q))
rows = W.c2_rows(qids, n_obs)
starts = np.tile(np.array(cfg["start"]), (Q, 1)); goals = np.tile(np.array(cfg["goal"]), (Q, 1))
batch = rrtk.RRTStarBatch(starts, goals, rows, cfg["rand_area"], cfg["expand_dis"], cfg["path_resolution"],
                          cfg["goal_sample_rate"], iters, None, cfg["robot_radius"], "sobol",
                          cfg["connect_circle_dist"], True, seed=0xC2,
                          sobol_offset=np.asarray(qids, dtype=np.int64) * iters)
import pynvml
pynvml.nvmlInit()
_h = pynvml.nvmlDeviceGetHandleByIndex(0)
def mhz():
    return pynvml.nvmlDeviceGetClockInfo(_h, pynvml.NVML_CLOCK_SM)
for k in range(4):
    t0 = time.perf_counter(); batch.run(); torch.cuda.synchronize()
    print("warm run", k, f"{(time.perf_counter() - t0) * 1e3:.1f} ms, sm {mhz()} MHz", flush=True)
t0 = time.perf_counter()
h_path = torch.empty((Q, 256, 2), dtype=torch.float64).pin_memory()
h_plen = torch.empty((Q,), dtype=torch.int32).pin_memory()
print(f"pin_memory {(time.perf_counter() - t0) * 1e3:.1f} ms, sm {mhz()} MHz", flush=True)
for k in range(3):
    t0 = time.perf_counter(); batch.run(); torch.cuda.synchronize()
    print("post-pin run", k, f"{(time.perf_counter() - t0) * 1e3:.1f} ms, sm {mhz()} MHz", flush=True)
def cpustat():
    try:
        d = dict(l.split() for l in open("/sys/fs/cgroup/cpu.stat"))
        return "thr=%s/%s us" % (d.get("nr_throttled"), d.get("throttled_usec"))
    except Exception as e:
        return "nocg"
print(open("/sys/fs/cgroup/cpu.max").read().strip() if os.path.exists("/sys/fs/cgroup/cpu.max") else "no cpu.max", "load", os.getloadavg(), flush=True)
def reasons():
    try:
        return hex(pynvml.nvmlDeviceGetCurrentClocksEventReasons(_h))
    except Exception:
        return hex(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(_h))
print("procs on gpu:", len(pynvml.nvmlDeviceGetComputeRunningProcesses(_h)), "cpu count", os.cpu_count(),
      "affinity", len(os.sched_getaffinity(0)), flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for k in range(40):
    t = [time.perf_counter()]
    batch.upload(); torch.cuda.synchronize(); t.append(time.perf_counter())
    e0.record(); tl = time.perf_counter(); r = batch.run(); tl = time.perf_counter() - tl; e1.record(); torch.cuda.synchronize(); t.append(time.perf_counter())
    print("launch call %.2f ms" % (tl * 1e3), cpustat(), end=" ")
    print("   kernel by events %.1f ms" % e0.elapsed_time(e1), reasons(), end=" | ")
    path, plen = r.paths_device(256); torch.cuda.synchronize(); t.append(time.perf_counter())
    h_path.copy_(path, non_blocking=True); h_plen.copy_(plen, non_blocking=True); torch.cuda.synchronize(); t.append(time.perf_counter())
    print("step", k, " ".join(f"{(b - a) * 1e3:.1f}" for a, b in zip(t, t[1:])), "ms (upload, run, paths, d2h)", mhz(), "MHz", flush=True)
