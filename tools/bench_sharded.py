"""BASELINE configs 4 and 5 sharded over the GPUs of one box: independent units, static block
partition, no data-path collective; NCCL only for the barrier, the max-over-ranks timing and the final gather.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/bench_sharded.py
"""
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "robotics-path-planning_b200"))
import numpy as np
import torch
import rrtk
from rrtk import arm as A, dist as D, dubins_planner as DP

rank, local, world = D.env_world()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
D.init("nccl", dev)
import torch.distributed as dist


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


out = {"n_gpus": world}
# ---- config 5: arm C-space grid M = 8192 x 64 obstacle sets, sharded by obstacle set ----
M, S = 8192, 64
rng = np.random.default_rng(5)
sets = np.concatenate([rng.uniform(-2, 2, (S, 5, 2)), rng.uniform(0.2, 0.7, (S, 5, 1))], axis=2)
lo, hi = rrtk.shard_range(S, rank, world)
for rep in range(2):
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    grid = A.occupancy_grids_device([0.5, 0.5, 0.3, 0.5, 0.1], sets[lo:hi], M, device=dev)
    b.record()
    barrier()
    t = D.max_over_ranks(a.elapsed_time(b) / 1e3, dev)
occ = grid.sum(dim=(1, 2)).to(torch.int64).reshape(-1, 1)          # per-set summary, gathered over NCCL
allocc = D.gather_summaries(occ) if (hi - lo) * world == S else occ
out["c5_arm_grid"] = dict(M=M, sets=S, sets_per_gpu=hi - lo, s=t, cells_per_s=M * M * S / t,
                          occupied_total=int(allocc.sum().item()))
del grid
# ---- config 4: RRT*-Dubins 1024 queries x 500 iterations, sharded by query ----
Q, iters = 1024, 500
rng = np.random.default_rng(7)
st = np.concatenate([rng.uniform(-2, 15, (Q, iters, 2)), rng.uniform(-math.pi, math.pi, (Q, iters, 1))], axis=2)
coin = rng.integers(0, 101, (Q, iters)) <= 10
st[coin] = (10.0, 10.0, 0.0)
obs = [(5, 5, 1), (3, 6, 2), (3, 8, 2), (3, 10, 2), (7, 5, 2), (9, 5, 2)]
lo, hi = rrtk.shard_range(Q, rank, world)
for rep in range(2):
    barrier()
    t0 = time.perf_counter()
    res = DP.run_batch([[0.0, 0.0, 0.0]] * (hi - lo), [[10.0, 10.0, 0.0]] * (hi - lo), [obs] * (hi - lo), 3.0, iters, st[lo:hi], device=dev)
    barrier()
    t = D.max_over_ranks(time.perf_counter() - t0, dev)
summ = torch.tensor([[r["n"], r["goal_index"]] for r in res], dtype=torch.int64, device=dev)
allsum = D.gather_summaries(summ)
out["c4_rrtstar_dubins"] = dict(queries=Q, iters=iters, queries_per_gpu=hi - lo, s_e2e=t, tree_iters_per_s_e2e=Q * iters / t,
                                solved=int((allsum[:, 1] >= 0).sum().item()), gathered_rows=int(allsum.shape[0]))
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
