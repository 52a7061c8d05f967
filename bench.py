#!/usr/bin/env python
"""bench.py -- RRT* tree-iterations/sec on BASELINE config 2 (4096 batched queries x 256 circles x
2000 iterations per GPU), plus the NN-search HBM roofline, next to the CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU); queries are partitioned statically across ranks
(weak scaling: 4096 queries per GPU), no data-path collective; NCCL is used for the barrier, the
max-over-ranks timing and the final gather of per-query result hashes.  A "step" is one pass of the hot
path over the whole batch: every query runs all 2000 iterations (search_until_max_iter=True).

Besides the weak-scaling headline the line carries, for every N, the configurations BASELINE.json shards
(`sharded`): config 2 at its own size (4096 queries in total, 4096 / N per GPU), config 4 (1024 RRT*-Dubins queries,
1024 / N per GPU) and config 5 (64 obstacle sets of the 8192^2 arm grid, 64 / N per GPU), each with its time, the
speed-up over the same work on one GPU (run on rank 0 in the same job) and a shard-invariance verdict: the 64-bit
result hash of every unit, all-gathered over NCCL, equals the unsharded run's.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "robotics-path-planning_b200"))

METRIC = "RRT* tree-iterations/sec (4096 batched queries x 256 circles x 2000 iterations per GPU)"
UNIT = "tree-iterations/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="rrtk", choices=["rrtk", "reference"])
    ap.add_argument("--queries", type=int, default=4096, help="queries per GPU")
    ap.add_argument("--iters", type=int, default=2000)
    ap.add_argument("--obstacles", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nn-nodes", type=int, default=1 << 26, help="nodes in the NN-search roofline run")
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configs (arm grid, Dubins, informed)")
    ap.add_argument("--no-sharded", action="store_true", help="skip the strong-scaling / sharded curves (configs 2, 4, 5)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port (pure Python, the reference's own arithmetic and data structures'
# cost profile) on a bounded, stratified sample of the same workload
# ------------------------------------------------------------------------------------------------
def _cpu_worker(job):
    """One query: fast-forward the tree with the C oracle to each stratum start, then time the Python
    port for `span` iterations from there.  Returns the estimated seconds for the full query."""
    qid, iters, n_obs, strata, span, stream = job
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "robotics-path-planning_b200"))
    import numpy as np
    import oracle as O
    import pyport
    from rrtk import workloads as W
    cfg = W.C2
    obs = W.c2_obstacles(qid, n_obs)
    obs_list = [tuple(r) for r in obs.tolist()]
    est = 0.0
    timed_iters = 0
    timed_s = 0.0
    for mid in strata:
        s0 = max(0, mid - span // 2)
        p, ob = O.make_params(cfg["start"], cfg["goal"], obs_list, cfg["expand_dis"], cfg["path_resolution"],
                              s0, None, cfg["robot_radius"], cfg["connect_circle_dist"], True,
                              math_mode=O.MATH_LIBM)
        pre = O.rrtstar_run(p, ob, stream[:max(s0, 1)], want_trace=False) if s0 > 0 else None
        port = pyport.RRTStarPort(cfg["start"], cfg["goal"], obs_list, cfg["expand_dis"],
                                  cfg["path_resolution"], s0 + span, None, cfg["robot_radius"],
                                  cfg["connect_circle_dist"], True)
        if pre is not None:
            port.x, port.y = pre["x"].tolist(), pre["y"].tolist()
            port.cost, port.parent = pre["cost"].tolist(), [int(v) for v in pre["parent"]]
        t0 = time.perf_counter()
        port.planning([tuple(r) for r in stream[:s0 + span]], first_iter=s0)
        dt = time.perf_counter() - t0
        timed_iters += span
        timed_s += dt
        est += dt / span * (iters / len(strata))
    return est, timed_iters, timed_s


def cpu_baseline(iters, n_obs, budget_s=20.0, cores=None, span=None):
    import multiprocessing as mp
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    O.build()
    from rrtk import sampling
    cores = cores or max(1, len(os.sched_getaffinity(0)))
    n_strata = 10
    # stratum k stands for iterations [k, k+1) * iters / n_strata and is sampled around its midpoint
    strata = [int(iters * (k + 0.5) / n_strata) for k in range(n_strata)]
    # ~45 ms per iteration at 1500 nodes / 256 obstacles in CPython: size the spans to the budget
    span = span or max(4, int(budget_s / n_strata / 0.030))
    span = min(span, iters // n_strata // 2)
    jobs = []
    for w in range(cores):
        # same stream the in-kernel sampler draws for query w (host restatement of the coins + oracle Sobol)
        coins = sampling.kernel_coins(0xC2, w, iters, 5)
        pts = O.sobol_fill(2, w * iters, int((~coins).sum()))
        stream = np.empty((iters, 2))
        stream[~coins] = -2.0 + pts * 17.0
        stream[coins] = (13.0, 13.0)
        jobs.append((w, iters, n_obs, strata, span, stream))
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    rate = sum(iters / est for est, _, _ in res)
    return dict(value=rate, unit=UNIT, cores=cores, kind="port",
                sample=f"{cores} queries (one per core) x {n_strata} strata x {span} iterations of the "
                       f"{n_obs}-obstacle / {iters}-iteration workload in the pure-Python port "
                       f"(oracle/pyport.py); tree state at each stratum start fast-forwarded by the C oracle; "
                       f"rate extrapolated to full queries; wall {wall:.1f} s",
                per_core=rate / cores)


def cpu_baseline_c(iters, n_obs, cores=None, per_core=2):
    """The C restatement (oracle/rrtk_oracle.c, libm mode), full queries, one thread per core."""
    from concurrent.futures import ThreadPoolExecutor
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    from rrtk import sampling, workloads as W
    cfg = W.C2
    cores = cores or max(1, len(os.sched_getaffinity(0)))
    jobs = []
    for w in range(cores * per_core):
        coins = sampling.kernel_coins(0xC2, w, iters, 5)
        pts = O.sobol_fill(2, w * iters, int((~coins).sum()))
        stream = np.empty((iters, 2))
        stream[~coins] = -2.0 + pts * 17.0
        stream[coins] = (13.0, 13.0)
        p, ob = O.make_params(cfg["start"], cfg["goal"], W.c2_obstacles(w, n_obs).tolist(), cfg["expand_dis"],
                              cfg["path_resolution"], iters, None, 0.0, cfg["connect_circle_dist"], True)
        jobs.append((p, ob, stream))
    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        list(ex.map(lambda j: O.rrtstar_run(j[0], j[1], j[2], want_trace=False), jobs))
    wall = time.perf_counter() - t0
    return dict(value=len(jobs) * iters / wall, unit=UNIT, cores=cores, kind="port-c",
                sample=f"{len(jobs)} full queries in the C oracle (gcc -O2, FP64, brute force), {cores} threads")


# ------------------------------------------------------------------------------------------------
def _cpu_full_query(job):
    """One whole query of the workload in the pure-Python port (oracle/pyport.py): every iteration is executed and timed."""
    qid, iters, n_obs, stream = job
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "robotics-path-planning_b200"))
    import pyport
    from rrtk import workloads as W
    cfg = W.C2
    obs_list = [tuple(r) for r in W.c2_obstacles(qid, n_obs).tolist()]
    port = pyport.RRTStarPort(cfg["start"], cfg["goal"], obs_list, cfg["expand_dis"], cfg["path_resolution"], iters, None,
                              cfg["robot_radius"], cfg["connect_circle_dist"], True)
    t0 = time.perf_counter()
    port.planning([tuple(r) for r in stream[:iters]])
    return time.perf_counter() - t0, len(port.x)


def _c2_streams(n, iters):
    """The sample stream the in-kernel sampler draws for queries 0 .. n-1 (host restatement of the coins + oracle Sobol)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    from rrtk import sampling
    out = []
    for w in range(n):
        coins = sampling.kernel_coins(0xC2, w, iters, 5)
        pts = O.sobol_fill(2, w * iters, int((~coins).sum()))
        stream = np.empty((iters, 2))
        stream[~coins] = -2.0 + pts * 17.0
        stream[coins] = (13.0, 13.0)
        out.append(stream)
    return out


def reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The reference is pure Python
    and cannot travel to the GPU box, so this times the oracle port (oracle/pyport.py: same arithmetic, same O(n) list
    scans per iteration, bit-identical results).  A step = one WHOLE query per host core, all `iters` iterations, timed
    by the wall clock around the pool (nothing is extrapolated); warm-up steps run 100-iteration queries."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    O.build()
    cores = max(1, len(os.sched_getaffinity(0)))
    streams = _c2_streams(cores, args.iters)
    t_all = time.perf_counter()
    walls, nodes = [], []
    with mp.get_context("fork").Pool(cores) as pool:
        for k in range(args.warmup + args.steps):
            it = min(100, args.iters) if k < args.warmup else args.iters
            jobs = [(w, it, args.obstacles, streams[w]) for w in range(cores)]
            t0 = time.perf_counter()
            res = pool.map(_cpu_full_query, jobs)
            if k >= args.warmup:
                walls.append(time.perf_counter() - t0)
                nodes.append(sum(r[1] for r in res) / cores)
    per_step = cores * args.iters
    v = per_step * len(walls) / sum(walls)
    cb = dict(value=v, unit=UNIT, cores=cores, kind="port", per_core=v / cores,
              sample=f"{cores} whole queries per step (one per host core), all {args.iters} iterations of the "
                     f"{args.obstacles}-obstacle workload in the pure-Python port (oracle/pyport.py), wall clock around "
                     f"the process pool; {len(walls)} timed steps of {min(walls):.1f}-{max(walls):.1f} s; mean tree "
                     f"{sum(nodes) / len(nodes):.0f} nodes")
    out = dict(metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
               ms_per_step=1e3 * sum(walls) / len(walls), higher_is_better=True, scaling="weak",
               vs_baseline=None, dtype="f64", data="synthetic", impl="reference",
               config=dict(workload="c2: batched RRT* (rrt_04 semantics), 256 random circles, 2000 iterations, "
                                    "expand_dis 1.0 / resolution 0.1, Sobol sampler",
                           queries_per_step=cores, obstacles=args.obstacles, iters=args.iters),
               cpu_baseline=cb,
               e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
               wall_s=time.perf_counter() - t_all)
    _emit(json.dumps(out))


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.first = index, None, [], 0

    def mark(self):
        """The samples from here on are the ones reported (call at the start of the timed region)."""
        self.first = len(self.lines)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[self.first:] or self.lines[-1:]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


def _emit(line: str) -> None:
    """Write the result line to the process's ORIGINAL stdout (see main)."""
    os.write(_REAL_STDOUT, (line + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    args = parse()
    # stdout carries exactly ONE line, the JSON result: libraries that print there (NCCL's version banner under
    # NCCL_DEBUG=VERSION, a child process) are sent to stderr for the whole run
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return reference_arm(args)
    import numpy as np
    import torch
    import rrtk
    from rrtk import _lib, workloads as W

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    L = _lib.lib()
    Q, iters, n_obs = args.queries, args.iters, args.obstacles
    cfg = W.C2

    # ---- this rank's shard of the global batch (weak scaling: Q queries per GPU) ----
    lo, hi = rrtk.shard_range(Q * world, rank, world)
    qids = list(range(lo, hi))
    rows = W.c2_rows(qids, n_obs)
    starts = np.tile(np.array(cfg["start"]), (len(qids), 1))
    goals = np.tile(np.array(cfg["goal"]), (len(qids), 1))
    batch = rrtk.RRTStarBatch(starts, goals, rows, cfg["rand_area"], cfg["expand_dis"], cfg["path_resolution"],
                              cfg["goal_sample_rate"], iters, None, cfg["robot_radius"], "sobol",
                              cfg["connect_circle_dist"], True, seed=0xC2, device=dev, query_base=lo)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # the clock sampler (one `nvidia-smi -lms 100` process) is started BEFORE the warm-up: its start-up attaches to the driver
    # and, when it fell inside the 0.25 s timed region, cost a launch up to 20 %; only the samples taken during the timed
    # region are reported
    clocks = ClockSampler(local)
    clocks.start()
    for _ in range(args.warmup):
        batch.run()
    barrier()

    # ---- device-timed region: K launches of the persistent kernel, inputs resident in HBM.  Every launch has its own
    # event pair (no host synchronisation in between): the step time is the MEDIAN launch, max over ranks; the mean of
    # the whole bracket and the min / max launch are reported beside it ----
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    clocks.mark()
    for a, b in evs:
        a.record()
        batch.run()
        b.record()
    barrier()
    launch_ms = sorted(a.elapsed_time(b) for a, b in evs)
    t_bracket = max_over_ranks(evs[0][0].elapsed_time(evs[-1][1]) / 1e3)
    t_med = max_over_ranks(launch_ms[len(launch_ms) // 2] / 1e3)
    t_dev = t_med * args.steps            # (the roofline and ms_per_step below are per median launch)
    launch_stats = dict(min=launch_ms[0], median=launch_ms[len(launch_ms) // 2], max=launch_ms[-1],
                        median_max_over_ranks=t_med * 1e3, mean_of_bracket=1e3 * t_bracket / args.steps,
                        timing="one CUDA event pair per launch on the launching stream; value uses the median launch")
    clk = clocks.stop()
    total_iters = float(world) * Q * iters * args.steps
    value = total_iters / t_dev
    res = batch.result
    status = res.status.cpu().numpy()
    n_nodes = res.n_nodes.cpu().numpy()
    assert (status == 0).all(), "a query overflowed its near list / node capacity"

    # ---- end to end: pinned host inputs -> H2D -> kernel -> path extraction -> D2H, every step ----
    path_cap = 256
    h_path = torch.empty((len(qids), path_cap, 2), dtype=torch.float64).pin_memory()
    h_plen = torch.empty((len(qids),), dtype=torch.int32).pin_memory()

    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def e2e_step():
        """One user-level call; returns (device ms from the first H2D byte to the last D2H byte, wall ms)."""
        w0 = time.perf_counter()
        ea.record()
        batch.upload()                                   # pinned host -> device (start/goal, obstacles, counts)
        r = batch.run()
        path, plen = r.paths_device(path_cap)            # generate_final_course on the device
        h_path.copy_(path, non_blocking=True)            # device -> pinned host: what planning() returns
        h_plen.copy_(plen, non_blocking=True)
        eb.record()
        eb.synchronize()                                 # the caller holds the paths in host memory here
        return ea.elapsed_time(eb), (time.perf_counter() - w0) * 1e3
    # the same call recorded once as a CUDA graph (RRTStarBatch.capture): one launch per step instead of ~10 enqueues, so
    # a descheduled host thread cannot open a gap between the copies and the kernel; falls back to the plain call
    graph_ok = False
    try:
        batch.capture(h_path, h_plen, path_cap)
        graph_ok = True
    except Exception as exc:  # noqa: BLE001
        sys.stderr.write(f"e2e: CUDA graph capture failed ({exc!r}); timing the plain call\n")
    if graph_ok:
        def e2e_step():  # noqa: F811
            w0 = time.perf_counter()
            ea.record()
            batch.replay()
            eb.record()
            eb.synchronize()
            return ea.elapsed_time(eb), (time.perf_counter() - w0) * 1e3
    for _ in range(max(3, args.warmup)):                 # untimed: allocator / pinned-buffer warm-up
        e2e_step()
    barrier()
    e2e_ms = [e2e_step() for _ in range(args.steps)]
    barrier()
    # timed on the device (CUDA events on the launching stream, copies inside), max over ranks; the wall clock of the
    # same calls is reported next to it -- on a freshly started box the host thread can be descheduled for
    # 0.1 - 2 s inside cudaStreamSynchronize while the device timeline stays at kernel + copies (tools/probe_e2e.py)
    t_e2e = max_over_ranks(sum(m[0] for m in e2e_ms) / 1e3)
    t_e2e_wall = max_over_ranks(sum(m[1] for m in e2e_ms) / 1e3)
    e2e = dict(value=total_iters / t_e2e, unit=UNIT, h2d_bytes_per_step=batch.h2d_bytes() * world,
               d2h_bytes_per_step=(h_path.numel() * 8 + h_plen.numel() * 4) * world,
               ms_per_step=1e3 * t_e2e / args.steps, timing="CUDA events around upload + kernel + path extraction + D2H",
               call="RRTStarBatch.replay() (the call captured as a CUDA graph)" if graph_ok else "RRTStarBatch.upload / run / paths_device",
               wall_value=total_iters / t_e2e_wall, wall_ms_per_step=1e3 * t_e2e_wall / args.steps,
               result="paths [Q, 256, 2] + lengths, pinned host buffers")
    found = int((h_plen.numpy() > 0).sum())
    e2e_d2h = h_path.numel() * 8 + h_plen.numel() * 4

    # ---- shard invariance of the headline batch: the 64-bit result hash of every query is all-gathered
    # over NCCL (8 B per query; not in the timed regions) and rank 0 re-plans a sample of query ids taken from EVERY
    # rank's block as one small unsharded batch: the hashes must be equal ----
    weak_inv = shard_invariance_sample(torch, dist, rrtk, W, dev, rank, world, Q, iters, n_obs, query_hash(torch, res))
    sharded = None
    if not args.no_sharded:
        del h_path, h_plen
        sharded = sharded_curves(torch, dist, rrtk, W, dev, rank, world, iters, n_obs, barrier, max_over_ranks)

    out = None
    if rank == 0:
        peaks = measured_peaks()
        ms_step = 1e3 * t_dev / args.steps

        # ---- pipe peaks measured here (FMA loops), NN-search HBM roofline ----
        def fma_peak(fp64):
            buf = torch.zeros(8, dtype=torch.float64, device=dev)
            blocks, it = 148 * 8, 20000
            for _ in range(2):
                L.rrtk_fma_peak_dev(fp64, it, blocks, buf.data_ptr(), torch.cuda.current_stream().cuda_stream)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            L.rrtk_fma_peak_dev(fp64, it, blocks, buf.data_ptr(), torch.cuda.current_stream().cuda_stream)
            b.record()
            torch.cuda.synchronize()
            return blocks * 256 * 8 * 2.0 * it / (a.elapsed_time(b) / 1e3) / 1e12
        fp64_peak, fp32_peak = fma_peak(1), fma_peak(0)

        nn = nn_roofline(torch, L, dev, args.nn_nodes, peaks)
        # SURVEY 8d's second size: 10^6 nodes (8 MB, L2-resident -- launch- and L2-bound, reported as a detail only)
        nn["detail"]["n_1e6"] = nn_roofline(torch, L, dev, 1_000_000, peaks)["detail"]

        # algorithmic work of one launch (DESIGN.md "Roofline"): the reference's brute-force FP64
        # point-circle tests and node-distance evaluations (counted offline, see ALGORITHMIC_FLOP_PER_ITER)
        alg = algorithmic_work(iters, n_obs)
        achieved = alg["flop_per_iter"] * Q * iters / (t_dev / args.steps) / 1e12 if alg["flop_per_iter"] else None
        roofline = dict(kernel="rrtstar_kernel", bound="fp64", achieved=achieved, peak=fp64_peak, unit="TFLOP/s",
                        frac=achieved / fp64_peak if achieved else None,
                        # dram__bytes_read.sum + dram__bytes_write.sum of one launch of THIS build at the default workload
                        # (ncu --set full, profiles/r2k_rrtstar_kernel_ncu_summary.txt: 7.891 + 2.149 GB; bench.py cannot
                        # read hardware counters itself -- re-captured whenever the kernel changes)
                        traffic=10.040e9 if (Q, iters, n_obs) == (4096, 2000, 256) else None, traffic_unit="bytes/launch",
                        traffic_source="profiles/r2k_rrtstar_kernel_ncu_summary.txt",
                        peak_source="FMA-loop probe measured in this run (rrtk_fma_peak_dev); "
                                    "MEASURED_PEAKS.json has no FP64 figure",
                        algorithmic_flop_per_iter=alg["flop_per_iter"], algorithmic_note=alg["note"],
                        fp32_fma_peak_tflops=fp32_peak)
        out = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                   ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                   data="synthetic",
                   config=dict(workload="c2: batched RRT* (rrt_04 semantics), 4096 queries/GPU x 256 random "
                                        "circles x 2000 iterations, expand_dis 1.0 / resolution 0.1, in-kernel "
                                        "Sobol sampler, search_until_max_iter=True",
                               queries_per_gpu=Q, obstacles=n_obs, iters=iters, parallelism=f"queries/{world}",
                               l2="working set (trees 229 MB + obstacles 34 MB per GPU) exceeds the 126 MB L2; "
                                  "no explicit flush"),
                   clocks=clk, e2e=e2e, gpu_launches=args.steps, roofline=roofline, roofline_nn=nn,
                   launch_ms=launch_stats, exec_mode={0: "auto", 1: "warp", 2: "cta"}[batch.params.exec_mode],
                   shard_invariance=weak_inv, paths_found=found, mean_nodes=float(n_nodes.mean()))
        if sharded is not None:
            out["sharded"] = sharded
        if not args.no_extras:
            out["extras"] = extras(torch, dev)
            try:   # path_smoothing (rrt_04:1447-1479) of the 4096 final courses, on the device
                from rrtk import smoothing
                sm_iters = 1000
                sp, sl = res.paths_device(64 + sm_iters)
                obs3 = batch.obstacles[:, :, :3].contiguous()          # robot_radius = 0: column 2 is the size
                draws = torch.rand((len(qids), sm_iters, 2), dtype=torch.float64, device=dev)
                len0 = float(sl.double().mean().item())
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                st, _ = smoothing.smooth_batch(sp, sl, sm_iters, obs3, batch.n_obs, draws)
                b.record()
                torch.cuda.synchronize()
                out["extras"]["path_smoothing"] = dict(paths=len(qids), iters=sm_iters, ms=a.elapsed_time(b),
                                                       mean_points_before=len0, mean_points_after=float(sl.double().mean().item()),
                                                       failed=int((st != 0).sum().item()))
            except Exception as e:  # noqa: BLE001
                out["extras"]["path_smoothing"] = dict(error=repr(e))
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(iters, n_obs)
            out["cpu_baseline_c"] = cpu_baseline_c(iters, n_obs)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        _emit(json.dumps(out))


def query_hash(torch, res):
    """64-bit hash per query of (n_nodes, goal_index, bits of cost[goal_index]) -- [Q] int64 on the device."""
    gi = res.goal_index.long()
    cg = res.cost.gather(1, gi.clamp(min=0).unsqueeze(1)).squeeze(1).contiguous().view(torch.int64)
    cg = torch.where(gi >= 0, cg, torch.zeros_like(cg))
    return (cg * 1000003) ^ (res.n_nodes.long() << 40) ^ ((gi + 1) << 20) ^ (res.status.long() << 60)


def all_gather_i64(torch, dist, t, world):
    """all_gather of equally sized int64 vectors in rank order (NCCL)."""
    if dist is None:
        return t
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t.contiguous())
    return torch.cat(parts, 0)


def _c2_batch(rrtk, W, dev, qids, iters, n_obs):
    import numpy as np
    cfg = W.C2
    rows = W.c2_rows(qids, n_obs)
    starts = np.tile(np.array(cfg["start"]), (len(qids), 1))
    goals = np.tile(np.array(cfg["goal"]), (len(qids), 1))
    return rows, starts, goals, cfg


def shard_invariance_sample(torch, dist, rrtk, W, dev, rank, world, Q, iters, n_obs, my_hash, n_sample=64):
    """Rank r planned queries [r * Q, (r + 1) * Q).  Gather every query's hash; rank 0 re-plans `n_sample` query ids spread
    over all ranks' blocks as ONE batch (different batch composition, different launch geometry) and compares."""
    import numpy as np
    allh = all_gather_i64(torch, dist, my_hash, world)
    out = None
    if rank == 0:
        total = Q * world
        ids = np.unique(np.linspace(0, total - 1, min(n_sample, total)).astype(np.int64))
        rows, starts, goals, cfg = _c2_batch(rrtk, W, dev, ids.tolist(), iters, n_obs)
        b = rrtk.RRTStarBatch(starts, goals, rows, cfg["rand_area"], cfg["expand_dis"], cfg["path_resolution"],
                              cfg["goal_sample_rate"], iters, None, cfg["robot_radius"], "stream",
                              cfg["connect_circle_dist"], True, seed=0xC2, device=dev,
                              sample_stream=_streams_for(torch, rrtk, W, dev, ids, iters))
        h = query_hash(torch, b.run())
        same = h == allh[torch.from_numpy(ids).to(dev)]
        out = dict(queries_gathered=int(allh.numel()), gather_bytes=int(allh.numel() * 8), sample=int(len(ids)),
                   sample_equal=int(same.sum().item()), ok=bool(same.all().item()),
                   how="hash(n_nodes, goal_index, cost[goal], status) of every query all-gathered over NCCL; rank 0 "
                       "re-plans the sampled ids (taken from every rank's block) as one unsharded batch")
        del b
    return out


def _streams_for(torch, rrtk, W, dev, ids, iters):
    """The samples the in-kernel sampler draws for GLOBAL query ids `ids` (each materialised with query_base = id)."""
    import numpy as np
    from rrtk import engine, _lib
    cfg = W.C2
    out = torch.empty((len(ids), iters, 2), dtype=torch.float64, device=dev)
    sg = torch.tensor([[cfg["start"][0], cfg["start"][1], cfg["goal"][0], cfg["goal"][1]]], dtype=torch.float64, device=dev)
    for k, qid in enumerate(ids.tolist()):
        p = engine.make_params(1, iters, iters + 1, 1, cfg["expand_dis"], cfg["path_resolution"], None, True,
                               _lib.SAMPLER_SOBOL, cfg["goal_sample_rate"], cfg["rand_area"][0], cfg["rand_area"][1],
                               0xC2, query_base=qid)
        off = torch.tensor([qid * iters], dtype=torch.int64, device=dev)
        out[k] = engine.sample_stream_dev(p, sg, off)[0]
    return out.cpu().numpy()


def _median_ms(torch, fn, warm, reps):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ms.append(a.elapsed_time(b))
    ms.sort()
    return ms[len(ms) // 2], ms


def sharded_curves(torch, dist, rrtk, W, dev, rank, world, iters, n_obs, barrier, max_over_ranks):
    """The configurations BASELINE.json shards, at their own sizes (strong scaling): total work fixed, 1 / N of it per GPU,
    no data-path collective.  For each: median time of the shard launch (max over ranks), the same work unsharded on rank 0
    in this job (its time is the N = 1 point, its results the invariance reference), speed-up, and the hashes gathered
    over NCCL."""
    import math
    import numpy as np
    from rrtk import arm as A, dubins_planner as DP
    out = {}
    # ---- config 2: 4096 queries in total ----
    try:
        total = 4096
        lo, hi = rrtk.shard_range(total, rank, world)
        rows, starts, goals, cfg = _c2_batch(rrtk, W, dev, list(range(lo, hi)), iters, n_obs)
        mk = lambda r, s_, g, base: rrtk.RRTStarBatch(s_, g, r, cfg["rand_area"], cfg["expand_dis"], cfg["path_resolution"],  # noqa: E731
                                                      cfg["goal_sample_rate"], iters, None, cfg["robot_radius"], "sobol",
                                                      cfg["connect_circle_dist"], True, seed=0xC2, device=dev, query_base=base)
        b = mk(rows, starts, goals, lo)
        barrier()
        med, ms = _median_ms(torch, b.run, 2, 5)
        barrier()
        t_n = max_over_ranks(med)
        allh = all_gather_i64(torch, dist, query_hash(torch, b.result), world) if (hi - lo) * world == total else None
        mode = {0: "auto", 1: "warp", 2: "cta"}[b.params.exec_mode]
        del b
        if rank == 0:
            if world > 1:
                rows, starts, goals, cfg = _c2_batch(rrtk, W, dev, list(range(total)), iters, n_obs)
                full = mk(rows, starts, goals, 0)
                t_1, _ = _median_ms(torch, full.run, 2, 5)
                ok = bool((query_hash(torch, full.result) == allh).all().item()) if allh is not None else None
                del full
            else:
                t_1, ok = t_n, True
            out["c2_strong"] = dict(total_queries=total, queries_per_gpu=hi - lo, iters=iters, obstacles=n_obs, ms=t_n,
                                    launch_ms_rank0=ms, tree_iters_per_s=total * iters / (t_n / 1e3), ms_one_gpu=t_1,
                                    speedup_vs_one_gpu=t_1 / t_n, shard_invariant=ok, gather_bytes=8 * total, exec_mode=mode)
    except Exception as e:  # noqa: BLE001
        out["c2_strong"] = dict(error=repr(e))
    barrier()
    # ---- config 4: RRT*-Dubins, 1024 queries x 500 iterations in total ----
    try:
        total, it4 = 1024, 500
        rng = np.random.default_rng(7)
        st = np.concatenate([rng.uniform(-2, 15, (total, it4, 2)), rng.uniform(-math.pi, math.pi, (total, it4, 1))], axis=2)
        st[rng.integers(0, 101, (total, it4)) <= 10] = (10.0, 10.0, 0.0)
        obs = [(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2)]

        def run4(a, b_):
            tm, kms, res = {}, [], None
            for _ in range(6):          # first call is the warm-up
                res = DP.run_batch([[0.0, 0.0, 0.0]] * (b_ - a), [[10.0, 10.0, 0.0]] * (b_ - a), [obs] * (b_ - a), 3.0, it4,
                                   st[a:b_], device=dev, timing=tm)
                kms.append(tm["kernel_ms"])
            kms = sorted(kms[1:])
            m63 = (1 << 63) - 1       # (Python ints do not wrap: keep the hash inside int64)
            h = np.array([((r["n"] << 40) ^ ((r["goal_index"] + 1) << 20) ^
                           (int(np.float64(r["cost"][r["goal_index"]]).view(np.int64)) * 1000003 if r["goal_index"] >= 0 else 0)) & m63
                          for r in res], dtype=np.int64)
            return kms[len(kms) // 2], kms, torch.from_numpy(h).to(dev), int(sum(r["goal_index"] >= 0 for r in res))
        lo, hi = rrtk.shard_range(total, rank, world)
        barrier()
        med, kms, h, solved = run4(lo, hi)
        barrier()
        t_n = max_over_ranks(med)
        allh = all_gather_i64(torch, dist, h, world) if (hi - lo) * world == total else None
        if rank == 0:
            if world > 1:
                t_1, _, h1, solved = run4(0, total)
                ok = bool((h1 == allh).all().item()) if allh is not None else None
            else:
                t_1, ok = t_n, True
            out["c4_rrtstar_dubins"] = dict(total_queries=total, queries_per_gpu=hi - lo, iters=it4, kernel_ms=t_n,
                                            kernel_ms_rank0=kms, tree_iters_per_s=total * it4 / (t_n / 1e3),
                                            kernel_ms_one_gpu=t_1, speedup_vs_one_gpu=t_1 / t_n, shard_invariant=ok,
                                            gather_bytes=8 * total, solved=solved)
    except Exception as e:  # noqa: BLE001
        out["c4_rrtstar_dubins"] = dict(error=repr(e))
    barrier()
    # ---- config 5: arm C-space grid 8192 x 8192, 64 obstacle sets in total, sharded by grid ROWS (every rank computes its
    # rows of all 64 sets: the forward kinematics of a cell is shared by the sets, so splitting the sets would repeat it on
    # every rank -- measured 5.85 x at 8 GPUs that way) ----
    try:
        M, S = 8192, 64
        rng = np.random.default_rng(5)
        sets = np.concatenate([rng.uniform(-2, 2, (S, 5, 2)), rng.uniform(0.2, 0.7, (S, 5, 1))], axis=2)
        sets[0] = [[1.75, 0.75, 0.6], [0.55, 1.5, 0.5], [0, -1, 0.7], [0, -0.6, 0.4], [-1, 1., 0.3]]
        link = [0.5, 0.5, 0.3, 0.5, 0.1]
        lo, hi = rrtk.shard_range(M, rank, world)
        keep = {}
        # inputs staged once (the launch is ~1 ms: per-call uploads and a 4 GB allocation would be what is timed)
        L5, link5 = rrtk._lib.lib(), np.ascontiguousarray(link, dtype=np.float64)
        d_theta, d_sets = torch.from_numpy(A.theta_list(M)).to(dev), torch.from_numpy(np.ascontiguousarray(sets)).to(dev)
        bufs = {}

        def run5(a, b_):
            g = bufs.get(b_ - a)
            if g is None:
                g = bufs[b_ - a] = torch.empty((S, b_ - a, M), dtype=torch.uint8, device=dev)
            rrtk._lib.check(L5.rrtk_arm_grid_dev(M, d_theta.data_ptr(), a, b_ - a, 5, link5.ctypes.data, d_sets.data_ptr(), S, 5,
                                                 g.data_ptr(), torch.cuda.current_stream().cuda_stream), "rrtk_arm_grid_dev")
            keep["g"] = g
        barrier()
        med, ms = _median_ms(torch, lambda: run5(lo, hi), 1, 3)
        barrier()
        t_n = max_over_ranks(med)
        occ = keep.pop("g").sum(dim=(1, 2), dtype=torch.int64)          # occupied cells per set in this rank's rows
        allocc = all_gather_i64(torch, dist, occ, world) if (hi - lo) * world == M else None
        if rank == 0:
            if world > 1:
                t_1, _ = _median_ms(torch, lambda: run5(0, M), 1, 3)
                occ1 = keep.pop("g").sum(dim=(1, 2), dtype=torch.int64)
                ok = bool((occ1 == allocc.view(world, S).sum(0)).all().item()) if allocc is not None else None
            else:
                t_1, ok, occ1 = t_n, True, occ
            out["c5_arm_grid"] = dict(M=M, sets=S, rows_per_gpu=hi - lo, ms=t_n, ms_rank0=ms, cells_per_s=M * M * S / (t_n / 1e3),
                                      written_gbs=M * M * S / (t_n / 1e3) / 1e9,
                                      ms_one_gpu=t_1, speedup_vs_one_gpu=t_1 / t_n, shard_invariant=ok, gather_bytes=8 * S * world,
                                      occupied_total=int(occ1.sum().item()), occupied_set0=int(occ1[0].item()),
                                      sharding="grid rows (each rank: M / N rows x all 64 sets)")
        keep.clear()
        bufs.clear()
    except Exception as e:  # noqa: BLE001
        out["c5_arm_grid"] = dict(error=repr(e))
    barrier()
    return out


def _timed(torch, fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / 1e3 / reps


def extras(torch, dev):
    """The other BASELINE configs, each a short device-timed run (inputs resident; parity is in tests/)."""
    import math
    import numpy as np
    import ctypes as C
    from rrtk import _lib, arm as A, dubins, dubins_planner as DP, informed as INF, engine
    L = _lib.lib()
    s = torch.cuda.current_stream().cuda_stream
    out = {}
    try:   # large-tree mode (rrtk/large_tree.py): get_nearest_node_index of a 6.7e7-node tree INSIDE the planner loop
        import rrtk
        n0, iters = 1 << 26, 40
        gen = torch.Generator(device=dev).manual_seed(5)
        seed = torch.rand((n0, 2), dtype=torch.float64, device=dev, generator=gen) * 17.0 - 2.0
        par = torch.arange(-1, n0 - 1, dtype=torch.int32, device=dev)
        pl = rrtk.RRTLarge([0.0, 0.0], [60.0, 60.0], [(5, 5, 1), (3, 6, 2), (3, 8, 2), (7, 5, 2), (9, 5, 2)], [-2, 15], 1.0, 0.1, 5, 8,
                           None, 0.0, capacity=n0 + iters + 1, seed_xy=seed, seed_parent=par, device=dev)
        del seed, par
        st = np.random.default_rng(9).uniform(-2, 15, (iters, 2))
        pl.planning(st[:8])
        pl.tree.time_scans = True
        pl.tree.stats.update(scans=0, scan_ms=0.0, candidates=0, nearest_queries=0)
        pl.max_iter = iters - 8
        pl.planning(st[8:])
        ts = pl.tree.stats
        peaks = measured_peaks()
        peak = peaks["hbm_gbs"] if peaks else 6650.0
        gbs = 8.0 * n0 * ts["scans"] / (ts["scan_ms"] / 1e3) / 1e9
        out["large_tree_rrt"] = dict(nodes=int(pl.tree.n), planner_iterations=iters - 8, scans=ts["scans"],
                                     us_per_scan=1e3 * ts["scan_ms"] / ts["scans"], scan_gbs=gbs, frac_of_hbm_peak=gbs / peak,
                                     fp64_rechecks_per_query=ts["candidates"] / max(ts["nearest_queries"], 1),
                                     note="two streaming passes over the FP32 mirror per get_nearest_node_index (FP32 argmin, then the "
                                          "candidates inside the rounding band), FP64 re-check; CUDA events around each pass "
                                          "(memset + scan kernel + unpack) inside RRTLarge.planning")
        del pl
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        out["large_tree_rrt"] = dict(error=repr(e))
    try:   # config 5: arm C-space grid, M = 8192, 64 obstacle sets, script arm (arm02:298-304)
        M, S = 8192, 64
        rng = np.random.default_rng(5)
        sets = np.concatenate([rng.uniform(-2, 2, (S, 5, 2)), rng.uniform(0.2, 0.7, (S, 5, 1))], axis=2)
        sets[0] = [[1.75, 0.75, 0.6], [0.55, 1.5, 0.5], [0, -1, 0.7], [0, -0.6, 0.4], [-1, 1., 0.3]]
        link = np.array([0.5, 0.5, 0.3, 0.5, 0.1])
        theta = torch.from_numpy(A.theta_list(M)).to(dev)
        d_obs = torch.from_numpy(sets).to(dev)
        grid = torch.empty((S, M, M), dtype=torch.uint8, device=dev)
        t = _timed(torch, lambda: L.rrtk_arm_grid_dev(M, theta.data_ptr(), 0, M, 5, link.ctypes.data, d_obs.data_ptr(),
                                                      S, 5, grid.data_ptr(), s), reps=2)
        out["c5_arm_grid"] = dict(cells_per_s=M * M * S / t, ms=t * 1e3, M=M, sets=S, out_gb=M * M * S / 1e9,
                                  written_gbs=M * M * S / t / 1e9, occupied_set0=int(grid[0].sum().item()),
                                  kernel="arm_grid_rows_kernel (row rasteriser; bound: the M^2 S byte store stream)")
        t0 = _timed(torch, lambda: grid.fill_(1), reps=3)
        out["c5_arm_grid"]["plain_fill_gbs"] = M * M * S / t0 / 1e9       # the same bytes by torch's fill_: the store ceiling here
        tc = _timed(torch, lambda: L.rrtk_arm_grid_cells_dev(M, theta.data_ptr(), 0, M, 5, link.ctypes.data, d_obs.data_ptr(),
                                                             S, 5, grid.data_ptr(), s), reps=1)
        out["c5_arm_grid"]["cell_by_cell_ms"] = tc * 1e3                    # every cell in the reference's order (cross-check kernel)
        del grid
    except Exception as e:  # noqa: BLE001
        out["c5_arm_grid"] = dict(error=repr(e))
    try:   # astar_torus (arm02:113-184) over 64 occupancy grids of the 2-link arm, M = 512, grid -> route on the GPU
        M, S = 512, 64
        rng = np.random.default_rng(15)
        ang, rad = rng.uniform(0, 2 * math.pi, (S, 5)), rng.uniform(0.9, 2.0, (S, 5))
        sets = np.stack([rad * np.cos(ang), rad * np.sin(ang), rng.uniform(0.15, 0.45, (S, 5))], axis=2)
        grids = A.occupancy_grids_device([1.0, 1.0], sets, M)
        host = grids.cpu().numpy()
        st, gl = [], []
        for k in range(S):
            free = np.argwhere(host[k] == 0)
            st.append(free[rng.integers(len(free))]); gl.append(free[rng.integers(len(free))])
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        routes, rlen, expanded = A.astar_torus_batch(grids, np.array(st), np.array(gl))
        b.record()
        torch.cuda.synchronize()
        out["astar_torus"] = dict(M=M, queries=S, ms=a.elapsed_time(b), routes_found=int((rlen > 0).sum().item()),
                                  mean_route_cells=float(rlen.double().mean().item()),
                                  cells_expanded=int(expanded.sum().item()))
        del grids, routes
    except Exception as e:  # noqa: BLE001
        out["astar_torus"] = dict(error=repr(e))
    try:   # Dubins steering primitive: 262144 edges among 16 circles
        n = 1 << 18
        rng = np.random.default_rng(6)
        f = np.column_stack([rng.uniform(0, 12, (n, 2)), rng.uniform(-math.pi, math.pi, n)])
        tt = np.column_stack([f[:, 0:2] + rng.uniform(-4, 4, (n, 2)), rng.uniform(-math.pi, math.pi, n)])
        obs = np.zeros((1, 16, 4)); o = rng.uniform(0, 12, (16, 2)); r = rng.uniform(0.2, 0.8, 16)
        obs[0, :, 0:2] = o; obs[0, :, 2] = r; obs[0, :, 3] = r * r
        d_f, d_t, d_o = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (f, tt, obs))
        d_c = torch.tensor([16], dtype=torch.int32, device=dev)
        mode = torch.empty(n, dtype=torch.int32, device=dev); ln = torch.empty((n, 3), dtype=torch.float64, device=dev)
        end = torch.empty((n, 3), dtype=torch.float64, device=dev); npt = torch.empty(n, dtype=torch.int32, device=dev)
        fr = torch.empty(n, dtype=torch.uint8, device=dev)
        t = _timed(torch, lambda: L.rrtk_dubins_steer_dev(n, 1.0, 0.1, d_f.data_ptr(), d_t.data_ptr(), None, d_o.data_ptr(),
                                                          16, d_c.data_ptr(), mode.data_ptr(), ln.data_ptr(), end.data_ptr(),
                                                          npt.data_ptr(), fr.data_ptr(), None, 0, s))
        out["dubins_steer"] = dict(edges_per_s=n / t, ms=t * 1e3, edges=n,
                                   course_points_per_s=float(npt.sum().item()) / t)
    except Exception as e:  # noqa: BLE001
        out["dubins_steer"] = dict(error=repr(e))
    try:   # Reeds-Shepp steering primitive, 262144 edges among 16 circles
        from rrtk import _lib as LL
        n = 1 << 18
        rng = np.random.default_rng(16)
        f = np.column_stack([rng.uniform(0, 12, (n, 2)), rng.uniform(-math.pi, math.pi, n)])
        tt = np.column_stack([f[:, 0:2] + rng.uniform(-4, 4, (n, 2)), rng.uniform(-math.pi, math.pi, n)])
        obs = np.zeros((1, 16, 4)); o = rng.uniform(0, 12, (16, 2)); r = rng.uniform(0.2, 0.8, 16)
        obs[0, :, 0:2] = o; obs[0, :, 2] = r; obs[0, :, 3] = r * r
        d_f, d_t, d_o = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (f, tt, obs))
        d_c = torch.tensor([16], dtype=torch.int32, device=dev)
        ty = torch.empty((n, 5), dtype=torch.int32, device=dev); ln = torch.empty((n, 5), dtype=torch.float64, device=dev)
        Lt = torch.empty(n, dtype=torch.float64, device=dev); npa = torch.empty(n, dtype=torch.int32, device=dev)
        end = torch.empty((n, 3), dtype=torch.float64, device=dev); npt = torch.empty(n, dtype=torch.int32, device=dev)
        fr = torch.empty(n, dtype=torch.uint8, device=dev)
        t = _timed(torch, lambda: L.rrtk_reeds_shepp_steer_dev(n, 1.0, 0.1, d_f.data_ptr(), d_t.data_ptr(), None, d_o.data_ptr(),
                                                               16, d_c.data_ptr(), ty.data_ptr(), ln.data_ptr(), Lt.data_ptr(),
                                                               npa.data_ptr(), end.data_ptr(), npt.data_ptr(), fr.data_ptr(),
                                                               None, 0, s))
        out["reeds_shepp_steer"] = dict(edges_per_s=n / t, ms=t * 1e3, edges=n, course_points_per_s=float(npt.sum().item()) / t,
                                        words_per_s=48 * n / t)
    except Exception as e:  # noqa: BLE001
        out["reeds_shepp_steer"] = dict(error=repr(e))
    try:   # config 4: RRT*-Dubins, 1024 queries x 500 iterations, built-in scenario (rrt_05:1804-1859)
        Q, iters = 1024, 500
        rng = np.random.default_rng(7)
        st = np.concatenate([rng.uniform(-2, 15, (Q, iters, 2)), rng.uniform(-math.pi, math.pi, (Q, iters, 1))], axis=2)
        coin = rng.integers(0, 101, (Q, iters)) <= 10
        st[coin] = (10.0, 10.0, 0.0)
        obs = [[(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2)]] * Q
        tm, kms = {}, []
        for rep in range(6):                                  # one warm-up call, then the median of five
            t0 = time.perf_counter()
            res = DP.run_batch([[0.0, 0.0, 0.0]] * Q, [[10.0, 10.0, 0.0]] * Q, obs, 3.0, iters, st, timing=tm)
            torch.cuda.synchronize()
            t = time.perf_counter() - t0
            kms.append(tm["kernel_ms"])
        kms = sorted(kms[1:]); tm["kernel_ms"] = kms[len(kms) // 2]
        out["c4_rrtstar_dubins"] = dict(tree_iters_per_s=Q * iters / (tm["kernel_ms"] / 1e3), kernel_ms=tm["kernel_ms"], kernel_ms_runs=kms,
                                        tree_iters_per_s_e2e=Q * iters / t, s=t, queries=Q, iters=iters,
                                        mean_nodes=float(np.mean([r["n"] for r in res])),
                                        solved=int(sum(r["goal_index"] >= 0 for r in res)))
    except Exception as e:  # noqa: BLE001
        out["c4_rrtstar_dubins"] = dict(error=repr(e))
    try:   # rrt_03: RRT-Dubins (plain RRT, one whole-warp Dubins edge per iteration), the config-4 scenario and streams
        Q, iters = 1024, 500
        rng = np.random.default_rng(7)
        st = np.concatenate([rng.uniform(-2, 15, (Q, iters, 2)), rng.uniform(-math.pi, math.pi, (Q, iters, 1))], axis=2)
        st[rng.integers(0, 101, (Q, iters)) <= 10] = (10.0, 10.0, 0.0)
        obs = [[(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2)]] * Q
        tm, kms = {}, []
        for rep in range(4):
            res = DP.run_rrt_batch([[0.0, 0.0, 0.0]] * Q, [[10.0, 10.0, 0.0]] * Q, obs, iters, st, timing=tm)
            kms.append(tm["kernel_ms"])
        kms = sorted(kms[1:])
        out["rrt_dubins"] = dict(tree_iters_per_s=Q * iters / (kms[1] / 1e3), kernel_ms=kms[1], kernel_ms_runs=kms, queries=Q,
                                 iters=iters, mean_nodes=float(np.mean([r["n"] for r in res])),
                                 solved=int(sum(r["goal_index"] >= 0 for r in res)))
    except Exception as e:  # noqa: BLE001
        out["rrt_dubins"] = dict(error=repr(e))
    try:   # rrt_01: basic RRT (no near / choose_parent / rewire), 4096 queries x 2000 iterations of the config-2 scenes
        import rrtk
        from rrtk import workloads as W2
        cfg, Q, iters = W2.C2, 4096, 2000
        rows = W2.c2_rows(list(range(Q)), 256)
        b = rrtk.RRTStarBatch(np.tile(np.array(cfg["start"]), (Q, 1)), np.tile(np.array(cfg["goal"]), (Q, 1)), rows, cfg["rand_area"],
                              cfg["expand_dis"], cfg["path_resolution"], cfg["goal_sample_rate"], iters, None, cfg["robot_radius"],
                              "sobol", cfg["connect_circle_dist"], True, seed=0xC2, sobol_offset=np.arange(Q, dtype=np.int64) * iters)
        b.params.rrt_only = 1
        t = _timed(torch, lambda: b.run(), reps=3)
        r = b.run()
        out["rrt_basic"] = dict(tree_iters_per_s=float(r.iters_done.sum().item()) / t, kernel_ms=t * 1e3, queries=Q, max_iter=iters,
                                mean_iters=float(r.iters_done.double().mean().item()),
                                mean_nodes=float(r.n_nodes.double().mean().item()),
                                note="rrt_01 stops at the first goal connection: iterations actually run are counted")
        del b, r
    except Exception as e:  # noqa: BLE001
        out["rrt_basic"] = dict(error=repr(e))
    try:   # config 3: ONE Informed RRT* tree grown to 10^6 nodes (rrt_07 semantics, built-in scenario), whole GPU on it
        cap, iters = 1_000_001, 1_700_000
        rng = np.random.default_rng(9)
        free = rng.uniform(-2, 15, (iters, 2)); coin = rng.integers(0, 101, iters) <= 10; free[coin] = (6.0, 10.0)
        ball = rng.random((iters, 2))
        obs = [(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2), (8, 10, 1)]
        INF.near_table(cap)                                   # host table of near radii (cached), outside the timing
        d_free, d_ball = torch.from_numpy(free).to(dev), torch.from_numpy(ball).to(dev)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run = INF.run_tree([0.0, 0.0], [6.0, 10.0], obs, 0.5, iters, d_free, d_ball, node_cap=cap)
        b.record()
        torch.cuda.synchronize()
        t = a.elapsed_time(b) / 1e3
        i = run.info
        # algorithmic bytes: 16 B per node per iteration (nearest + near over the tree); n grows
        # linearly in the accepted iterations, so the sum over iterations is ~ 16 * iters * n_final / 2
        alg_bytes = 16.0 * i["iters_done"] * i["n_nodes"] / 2.0
        out["c3_informed_single_tree"] = dict(
            nodes=i["n_nodes"], iterations=i["iters_done"], s=t, us_per_iter=1e6 * t / max(1, i["iters_done"]),
            tree_iters_per_s=i["iters_done"] / t, c_best=i["c_best"], mean_near=i["total_hits"] / max(1, i["iters_done"]),
            reached_node_cap=bool(i["status"] & 2), scan_gbs_algorithmic=alg_bytes / t / 1e9,
            goal_events=i["goal_events"], resamples=i["resamples"], grid=i["grid"],
            # batched kernel: cycles[0] = batches, cycles[1..5] = CTA 0's clocks in pass A | exchange A | extend + cut + pass B +
            # candidates | exchange B + cut #2 | apply + goal
            batches=int(i["cycles"][0]), samples_per_batch=i["iters_done"] / max(1, i["cycles"][0]),
            cycles_per_batch=[round(c / max(1, i["cycles"][0])) for c in i["cycles"][1:]],
            note="one fused FP64 pass (16 B/node) per iteration over an L2-resident tree; latency-bound by the "
                 "grid-wide exchange and the serial leaf math, see DESIGN.md 5.5")
        del run, d_free, d_ball
    except Exception as e:  # noqa: BLE001
        out["c3_informed_single_tree"] = dict(error=repr(e))
    try:   # RRT*-Reeds-Shepp (rrt_06), 512 queries x 300 iterations, built-in scenario (rrt_06:2015-2083)
        from rrtk import rs_planner as RP
        Q, iters = 512, 300
        rng = np.random.default_rng(17)
        st = np.concatenate([rng.uniform(-2, 15, (Q, iters, 2)), rng.uniform(-math.pi, math.pi, (Q, iters, 1))], axis=2)
        obs = [[(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2), (8, 10, 1)]] * Q
        tm, kms = {}, []
        for rep in range(6):
            t0 = time.perf_counter()
            res = RP.run_batch([[0.0, 0.0, 0.0]] * Q, [[10.0, 9.0, 0.0]] * Q, obs, 3.0, iters, st, robot_radius=0.6,
                               curvature=2.0, step_size=0.1, timing=tm)
            torch.cuda.synchronize()
            t = time.perf_counter() - t0
            kms.append(tm["kernel_ms"])
        kms = sorted(kms[1:]); tm["kernel_ms"] = kms[len(kms) // 2]
        out["rrtstar_reeds_shepp"] = dict(tree_iters_per_s=Q * iters / (tm["kernel_ms"] / 1e3), kernel_ms=tm["kernel_ms"], kernel_ms_runs=kms,
                                          tree_iters_per_s_e2e=Q * iters / t, s=t, queries=Q, iters=iters,
                                          mean_nodes=float(np.mean([r["n"] for r in res])),
                                          solved=int(sum(r["goal_index"] >= 0 for r in res)))
    except Exception as e:  # noqa: BLE001
        out["rrtstar_reeds_shepp"] = dict(error=repr(e))
    try:   # Closed-loop RRT* (rrt_10): 256 RRT*-RS trees with Reeds-Shepp-length costs x 100 iterations
        #        (built-in scenario, rrt_10:1610-1660), then pure-pursuit tracking of EVERY goal-reaching course
        from rrtk import closed_loop as CL, rs_planner as RP
        Q, iters = 256, 100
        rng = np.random.default_rng(19)
        st = np.concatenate([rng.uniform(-2, 20, (Q, iters, 2)), rng.uniform(-math.pi, math.pi, (Q, iters, 1))], axis=2)
        obs1 = [(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2), (8, 10, 1)]
        start, goal = [0.0, 0.0, 0.0], [6.0, 9.0, math.radians(90.0)]
        tm, tf, kms, fms = {}, {}, [], []
        for rep in range(6):
            t0 = time.perf_counter()
            trees = RP.run_batch([start] * Q, [goal] * Q, [obs1] * Q, float("inf"), iters, st, curvature=1.0, step_size=0.2,
                                 near_cap=224, timing=tm, rs_cost=True)
            kms.append(tm["kernel_ms"])
            if 0 < rep < 5:
                continue                                      # (the host-side course extraction is only timed once warm)
            courses = []
            for tr in trees:
                gi = [i for i in range(tr["n"]) if math.hypot(tr["x"][i] - goal[0], tr["y"][i] - goal[1]) <= 0.5
                      and abs(tr["yaw"][i] - goal[2]) <= math.radians(3.0)]
                courses += [np.asarray(c)[::-1] for c in CL.final_courses(tr, gi[:8], start, goal, 1.0, 0.2)]
            res = CL.closed_loop_batch(courses, obs1, timing=tf)
            t = time.perf_counter() - t0
            fms.append(tf["kernel_ms"])
        kms = sorted(kms[1:]); tm["kernel_ms"] = kms[len(kms) // 2]
        steps = int(sum(len(r["traj"]) for r in res))
        out["closed_loop_rrtstar"] = dict(
            queries=Q, iters=iters, planner_kernel_ms=tm["kernel_ms"], planner_kernel_ms_runs=kms, planner_tree_iters_per_s=Q * iters / (tm["kernel_ms"] / 1e3),
            courses=len(courses), feasible=int(sum(r["bits"] == 0 for r in res)), filter_kernel_ms=tf["kernel_ms"],
            tracking_steps_per_s=steps / (tf["kernel_ms"] / 1e3), s_e2e=t)
    except Exception as e:  # noqa: BLE001
        out["closed_loop_rrtstar"] = dict(error=repr(e))
    try:   # BIT* (rrt_08), 1024 queries x 200 counted iterations, built-in scenario (rrt_08:644-679)
        from rrtk import bitstar as BS
        Q, iters = 1024, 200
        rng = np.random.default_rng(23)
        draws = rng.random((Q, 6000))
        obs1 = [(5, 5, 0.5), (9, 6, 1), (7, 5, 1), (1, 5, 1), (3, 6, 1), (7, 9, 1)]
        tm, kms = {}, []
        for rep in range(6):
            t0 = time.perf_counter()
            res = BS.run_batch([[-1.0, 0.0]] * Q, [[3.0, 8.0]] * Q, [obs1] * Q, [-2, 15], iters, draws, timing=tm)
            t = time.perf_counter() - t0
            kms.append(tm["kernel_ms"])
        kms = sorted(kms[1:]); tm["kernel_ms"] = kms[len(kms) // 2]
        ok = [r for r in res if r["status"] == 0]
        out["bitstar"] = dict(queries=Q, iters=iters, kernel_ms=tm["kernel_ms"], kernel_ms_runs=kms, iters_per_s=Q * iters / (tm["kernel_ms"] / 1e3),
                              s_e2e=t, solved=int(sum(r["path_len"] > 0 for r in ok)), failed_status=Q - len(ok),
                              mean_batches=float(np.mean([r["batches"] for r in ok])),
                              edges_scored=float(np.mean([r["skipped"] + iters for r in ok])),
                              mean_expansions=float(np.mean([r["expansions"] for r in ok])))
    except Exception as e:  # noqa: BLE001
        out["bitstar"] = dict(error=repr(e))
    try:   # Informed RRT* (rrt_07 semantics), 512 queries x 1000 iterations, built-in scenario
        Q, iters = 512, 1000
        rng = np.random.default_rng(8)
        free = rng.uniform(-2, 15, (Q, iters, 2)); coin = rng.integers(0, 101, (Q, iters)) <= 10
        free[coin] = (6.0, 10.0)
        ball = rng.random((Q, iters, 2))
        obs = [[(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2), (8, 10, 1)]] * Q
        tm = {}
        for rep in range(2):
            t0 = time.perf_counter()
            res = INF.run_batch([[0.0, 0.0]] * Q, [[6.0, 10.0]] * Q, obs, 0.5, iters, free, ball, timing=tm)
            torch.cuda.synchronize()
            t = time.perf_counter() - t0
        out["informed_rrtstar"] = dict(tree_iters_per_s=Q * iters / (tm["kernel_ms"] / 1e3), kernel_ms=tm["kernel_ms"],
                                       tree_iters_per_s_e2e=Q * iters / t, s=t, queries=Q, iters=iters,
                                       mean_nodes=float(np.mean([r["n"] for r in res])),
                                       solved=int(sum(r["path"] is not None for r in res)))
    except Exception as e:  # noqa: BLE001
        out["informed_rrtstar"] = dict(error=repr(e))
    return out


# Reference-equivalent FP64 work per tree-iteration of the default workload (config 2: 256 circles, 2000 iterations):
# every node-distance evaluation of nearest / near (5 flop) and every point-circle test the reference's check_collision
# performs (5 flop), counted on 4 sample queries of the workload by tools/count_algorithmic_work.py (which runs the CPU
# oracle; bench.py's GPU arm does not): 28552.29 tests + 1274.86 distances per iteration.
ALGORITHMIC_FLOP_PER_ITER = {(2000, 256): 149135.72875}


def algorithmic_work(iters, n_obs):
    f = ALGORITHMIC_FLOP_PER_ITER.get((iters, n_obs))
    return dict(flop_per_iter=f,
                note="5 flop x (point-circle tests + node-distance evaluations) of the brute-force reference algorithm: "
                     "28552 tests + 1275 distances per iteration, counted offline on 4 sample queries "
                     "(tools/count_algorithmic_work.py)" if f else "no work count for this workload (run "
                     "tools/count_algorithmic_work.py)")


def nn_roofline(torch, L, dev, n, peaks):
    """NN-search kernel on an HBM-resident float2 array larger than L2 (8 bytes per node per pass)."""
    xy = torch.rand((n, 2), dtype=torch.float32, device=dev) * 17.0 - 2.0
    out = {}
    s = torch.cuda.current_stream().cuda_stream
    for B in (1, 8):
        smp = torch.rand((B, 2), dtype=torch.float32, device=dev) * 17.0 - 2.0
        scratch = torch.empty(B, dtype=torch.int64, device=dev)
        idx = torch.empty(B, dtype=torch.int32, device=dev)
        d2 = torch.empty(B, dtype=torch.float32, device=dev)
        for _ in range(3):
            _lib_check(L.rrtk_nearest_f32_dev(xy.data_ptr(), n, smp.data_ptr(), B, scratch.data_ptr(),
                                              idx.data_ptr(), d2.data_ptr(), s), L)
        reps = 10
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            L.rrtk_nearest_f32_dev(xy.data_ptr(), n, smp.data_ptr(), B, scratch.data_ptr(), idx.data_ptr(),
                                   d2.data_ptr(), s)
        b.record()
        torch.cuda.synchronize()
        dt = a.elapsed_time(b) / 1e3 / reps
        # check against torch on the same data
        ref = ((xy[None, :, :] - smp[:, None, :]) ** 2).sum(-1).argmin(1) if n <= (1 << 24) else None
        out[f"B{B}"] = dict(gbs=8.0 * n / dt / 1e9, ms=dt * 1e3)
        if ref is not None:
            out[f"B{B}"]["matches_torch_argmin"] = bool((ref.to(torch.int32) == idx).all().item())
    peak = peaks["hbm_gbs"] if peaks else 6650.0
    best = out["B1"]["gbs"]
    return dict(kernel="nearest_kernel<1>", bound="hbm", achieved=best, peak=peak, unit="GB/s", frac=best / peak,
                # ncu (profiles/r1_secondary_kernels_ncu_summary.txt): 536.9 MB read + 6.4 MB written per launch at
                # n = 2^26, i.e. 1.012 x the algorithmic 8 * n bytes
                traffic=543.24e6 if n == (1 << 26) else None, traffic_unit="bytes/launch", nodes=n, bytes_per_node=8,
                peak_source="MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s", detail=out)


def _lib_check(rc, L):
    if rc != 0:
        raise RuntimeError(L.rrtk_last_error().decode())


if __name__ == "__main__":
    main()
