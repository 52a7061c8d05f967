"""N > 1 host logic on CPU: world_size-2 gloo processes exercise the shard partition, the
max-over-ranks timing reduce and the final gather / min-cost reduce used by bench.py."""
import os
import subprocess
import sys
import textwrap

from conftest import ROOT

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, os.path.join(%r, "robotics-path-planning_b200"))
    import torch
    from rrtk import dist as D, shard_range
    rank, local, world = D.init("gloo")
    assert world == 2
    n = 11                                    # ragged on purpose
    lo, hi = D.my_shard(n)
    assert (lo, hi) == shard_range(n, rank, world)
    # every query id is owned exactly once
    own = torch.zeros(n, dtype=torch.int64); own[lo:hi] = 1
    import torch.distributed as dist
    dist.all_reduce(own)
    assert own.tolist() == [1] * n
    # timing: max over ranks
    assert D.max_over_ranks(1.0 + rank) == 2.0
    # final gather keeps global query order (equal shard sizes, as in bench.py: Q per GPU)
    q = 4
    summ = torch.stack([torch.arange(q) + rank * q, torch.full((q,), rank)], 1)
    g = D.gather_summaries(summ)
    assert g[:, 0].tolist() == list(range(world * q)) and g[:, 1].tolist() == [0] * q + [1] * q
    # replicas racing: min-cost reduce
    c = torch.tensor([3.0, 1.0]) if rank == 0 else torch.tensor([2.0, 5.0])
    assert D.best_of_replicas(c).tolist() == [2.0, 1.0]
    dist.barrier(); dist.destroy_process_group()
    print("rank", rank, "ok")
""") % ROOT


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29591", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {r} ok" in o
