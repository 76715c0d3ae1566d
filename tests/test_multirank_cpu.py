"""-m "not gpu": the N>1 host logic on CPU with the gloo backend, world_size 2 -- episode sharding is invariant
(rank r's shard of the synthetic workload equals the same global episodes generated in one process), the oracle
plans of the two shards concatenate to the single-process result, and the timing / counter reductions work."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import cpu_oracle as O
    from rl_mpc_lanemerging_b200 import sharding, synthetic
    lo, hi = sharding.shard_range(total, rank, world)
    S = synthetic.make_states(hi - lo, "moderate", seed=5, first_episode=lo, kind="mixed")
    r = O.plan_batch(O.default_params(), S["ego"], S["cars_x"], S["cars_v"], S["cars_a"], S["n_cars"], 18, layered=True, nthreads=2)
    mx = sharding.reduce_max([1.0 + rank, 5.0 - rank], "cpu")
    sm = sharding.reduce_sum([float(hi - lo)], "cpu")
    ga = sharding.gather([float(rank), 10.0 * rank], "cpu")       # the bench's per-rank diagnostics
    q.put((rank, lo, hi, r["idx"], r["cost"], mx, sm, ga))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(oracle):
    from rl_mpc_lanemerging_b200 import sharding, synthetic
    total, world = 37, 2                                   # odd on purpose: uneven shards
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [sharding.shard_range(total, r, world) for r in range(world)] == [(0, 19), (19, 37)]
    S = synthetic.make_states(total, "moderate", seed=5, kind="mixed")
    ref = oracle.plan_batch(oracle.default_params(), S["ego"], S["cars_x"], S["cars_v"], S["cars_a"], S["n_cars"], 18, layered=True)
    assert np.array_equal(np.concatenate([r[3] for r in res]), ref["idx"])
    assert np.array_equal(np.concatenate([r[4] for r in res]), ref["cost"])
    for r in res:
        assert r[5] == [2.0, 5.0] and r[6] == [float(total)]
        assert r[7] == [[0.0, 0.0], [1.0, 10.0]]
