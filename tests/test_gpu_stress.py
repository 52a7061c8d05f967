"""Randomised parity at volume: many random scenes in ONE launch of the production (non-trace) kernels, under both
executions of the loop, against the C oracle replayed query by query in threads.  RRTK_STRESS (default 24) sets the number
of scenes; profiles/r2_stress_parity.txt keeps the log of a run with RRTK_STRESS=384."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["warp", "cta"])
@pytest.mark.parametrize("expand,res,rr", [(1.0, 0.1, 0.0), (2.0, 0.25, 0.15)])
def test_random_scenes_in_one_launch_equal_the_oracle(mode, expand, res, rr, oracle_lib):
    import torch
    import rrtk
    from rrtk import _lib
    O = oracle_lib
    Q = int(os.environ.get("RRTK_STRESS", "24"))
    iters, o_max = 500, 96
    rng = np.random.default_rng(int(expand * 1000) + Q)
    lists, streams = [], np.empty((Q, iters, 2))
    for q in range(Q):
        n_obs = int(rng.integers(0, o_max + 1))
        obs = []
        while len(obs) < n_obs:
            x, y = rng.uniform(-2, 15, 2)
            r = rng.uniform(0.1, 0.6)
            if min(np.hypot(x, y), np.hypot(x - 13, y - 13)) > r + rr + 0.5:
                obs.append((float(x), float(y), float(r)))
        lists.append(obs)
        s = rng.uniform(-2, 15, (iters, 2))
        s[rng.integers(0, 101, iters) <= 5] = (13.0, 13.0)
        if q % 7 == 3:
            s[40:60] = s[39]                       # repeated samples: coincident nodes, equal d^2 (the `.index()` quirk)
        streams[q] = s
    starts = np.zeros((Q, 2)); goals = np.full((Q, 2), 13.0)
    b = rrtk.RRTStarBatch(starts, goals, lists, [-2, 15], expand, res, 5, iters, None, rr, "stream", 50.0, True,
                          near_cap=256, sample_stream=streams,
                          exec_mode={"warp": _lib.EXEC_WARP, "cta": _lib.EXEC_CTA}[mode])
    r = b.run()
    torch.cuda.synchronize()
    n_nodes = r.n_nodes.cpu().numpy(); par = r.parent.cpu().numpy(); xy = r.xy.cpu().numpy(); cost = r.cost.cpu().numpy()
    gi = r.goal_index.cpu().numpy(); status = r.status.cpu().numpy()

    def replay(q):
        p, obs = O.make_params([0.0, 0.0], [13.0, 13.0], lists[q], expand, res, iters, None, rr, 50.0, True, math_mode=O.MATH_CR)
        return O.rrtstar_run(p, obs if len(lists[q]) else np.zeros((0, 3)), streams[q], want_trace=False)
    with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 8)) as ex:
        refs = list(ex.map(replay, range(Q)))
    total = 0
    for q, ref in enumerate(refs):
        assert status[q] == 0, (q, status[q])
        n = ref["n"]
        assert n_nodes[q] == n, q
        assert np.array_equal(par[q, :n], ref["parent"]), q
        assert np.array_equal(xy[q, :n, 0], ref["x"]) and np.array_equal(xy[q, :n, 1], ref["y"]), q
        assert np.array_equal(cost[q, :n], ref["cost"]), q
        assert gi[q] == ref["goal_index"], q
        total += n
    print(f"stress {mode} expand={expand} res={res} rr={rr}: {Q} scenes x {iters} iterations, {total} nodes, all bit-identical to the oracle")
