"""Multi-rank host logic on CPU with gloo, world_size 2: the subproblem shards
idx = k*world + rank partition the search space (union of the shards' results == the whole search),
the incumbent merge is a MIN over ranks, counters sum, times take the max.  The shard workers here
are the CPU oracle standing in for the GPUs; the GPU versions of the same checks are in
tests/test_multi_gpu.py."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, seeds, depth, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle_py as orc
    from tests import tnf_gen
    import bench
    from turbo_b200 import abi
    results = []
    for seed in seeds:
        pb = tnf_gen.search_instance(seed) if seed < 8 else tnf_gen.random_net(16, 9, 5000 + seed, lo=-4, hi=4)
        r = orc.solve(pb, depth=depth, rank=rank, world=world)
        obj = r["objective"] if r["has_solution"] else abi.POS_INF
        t = torch.tensor([obj], dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)                 # the incumbent merge
        st = r["stats"]
        sums, maxes = bench.reduce_over_ranks(dist, [st["nodes"], st["eps_solved_subproblems"] + st["eps_skipped_subproblems"],
                                                     int(r["exhaustive"])], [float(rank + 1)])
        results.append((int(t.item()), sums, maxes))
    if rank == 0:
        out.put(results)
    dist.barrier()
    dist.destroy_process_group()


def test_two_shards_cover_the_search_space():
    from oracle import oracle_py as orc
    from tests import tnf_gen
    from turbo_b200 import abi
    seeds = list(range(12))
    depth = 5
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, seeds, depth, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for seed, (obj, sums, maxes) in zip(seeds, results):
        pb = tnf_gen.search_instance(seed) if seed < 8 else tnf_gen.random_net(16, 9, 5000 + seed, lo=-4, hi=4)
        whole = orc.solve(pb, depth=depth)
        expect = whole["objective"] if whole["has_solution"] else abi.POS_INF
        assert obj == expect, seed
        assert sums[2] == 2.0                      # both shards exhaustive
        assert maxes[0] == 2.0                     # max over ranks
        # without bound sharing the shards visit at least the nodes needed; every subproblem is
        # accounted for exactly once across the two shards
        assert sums[1] >= (1 << depth), (seed, sums)


def test_shard_with_known_incumbent_prunes():
    from oracle import oracle_py as orc
    from tests import tnf_gen
    pb = tnf_gen.search_instance(3)
    whole = orc.solve(pb, depth=4)
    assert whole["has_solution"]
    # a shard that already knows the optimum cannot improve on it and explores fewer nodes
    a = orc.solve(pb, depth=4, rank=0, world=2)
    b = orc.solve(pb, depth=4, rank=0, world=2, initial_bound=whole["objective"])
    assert not b["has_solution"] and b["exhaustive"]
    assert b["stats"]["nodes"] <= a["stats"]["nodes"]
