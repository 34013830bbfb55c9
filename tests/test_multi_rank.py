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


# ---- the final gather: tb_result_pack on every rank, a gather, tb_result_reduce on rank 0 ------------------

def fake_packed_result(pb, r):
    """What tb_result_pack would emit for a shard result `r` (the oracle's here: no GPU on this machine)."""
    import ctypes as C
    import numpy as np
    from turbo_b200 import abi
    h = abi.TbResultHeader()
    h.magic = abi.RESULT_MAGIC
    h.nvars = pb.nvars
    h.obj_var = pb.obj_var
    h.has_solution = int(r["has_solution"])
    h.exhaustive = int(r["exhaustive"])
    h.objective = r["objective"] if r["has_solution"] and pb.obj_var >= 0 else abi.POS_INF
    h.t_best_ns = r["stats"]["timers_ns"][abi.TIMER_LATEST_BEST_OBJ_FOUND]
    for name, _ in abi.TbStats._fields_:
        if name == "timers_ns":
            for i, v in enumerate(r["stats"][name]):
                h.stats.timers_ns[i] = v
        else:
            setattr(h.stats, name, r["stats"][name])
    return np.concatenate([np.frombuffer(bytes(h), np.uint8), np.asarray(r["lb"], np.int32).view(np.uint8),
                           np.asarray(r["ub"], np.int32).view(np.uint8)])


def _gather_worker(rank, world, port, seeds, depth, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle_py as orc
    from tests import tnf_gen
    from turbo_b200 import engine
    results = []
    for seed in seeds:
        pb = tnf_gen.search_instance(seed)
        r = orc.solve(pb, depth=depth, rank=rank, world=world)
        mine = torch.from_numpy(fake_packed_result(pb, r).copy())
        got = [torch.zeros_like(mine) for _ in range(world)] if rank == 0 else None
        dist.gather(mine, got, dst=0)                      # (NCCL on the GPU box, bench.py; gloo here)
        if rank == 0:
            m = engine.result_reduce([g.numpy() for g in got])
            results.append((m["has_solution"], int(m["lb"][pb.obj_var]) if m["has_solution"] else None, m["exhaustive"],
                            m["stats"]["nodes"], m["stats"]["eps_solved_subproblems"] + m["stats"]["eps_skipped_subproblems"],
                            m["best_rank"], r["stats"]["nodes"]))
    if rank == 0:
        out.put(results)
    dist.barrier()
    dist.destroy_process_group()


def test_final_gather_reports_the_global_best():
    from oracle import oracle_py as orc
    from tests import tnf_gen
    seeds = list(range(8))
    depth = 4
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, seeds, depth, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for seed, (has, obj, exh, nodes, subs, best_rank, nodes0) in zip(seeds, results):
        pb = tnf_gen.search_instance(seed)
        whole = orc.solve(pb, depth=depth)
        assert has == whole["has_solution"] and obj == whole["objective"] and exh
        assert subs >= (1 << depth) and nodes > nodes0          # counters are sums over the ranks
        if has:
            # the store that comes back is the winning rank's, whichever rank that is
            shard = orc.solve(pb, depth=depth, rank=best_rank, world=2)
            assert shard["objective"] == obj


def test_result_reduce_rejects_garbage():
    import numpy as np
    from turbo_b200 import engine
    with pytest.raises(engine.TurboError):
        engine.result_reduce([np.zeros(4096, np.uint8)])
