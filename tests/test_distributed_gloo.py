"""world_size-2 gloo runs (CPU) of the multi-rank host logic: contiguous sharding of a batch of
independent OCPs / scenarios over ranks, per-rank solves, gather in problem order, ordered total cost.
Each rank solves its shard with the host emulation of the device source (the GPU is not needed to
check the plumbing); rank 0 compares the gathered result with a single-process solve of the whole batch.
"""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import HostEmulation, random_x0
    from multi_agent_solver_b200 import sharding

    emu = HostEmulation()
    x0 = random_x0(2, total, seed=123)  # LQR agents: cheap
    lo, hi = sharding.shard_bounds(total, rank, world)
    mine = emu.solve(2, x0[lo:hi], np.zeros((hi - lo, 10, 4)), 100, 1e-5)
    X = sharding.allgather_rows(mine["X"], total)
    U = sharding.allgather_rows(mine["U"], total)
    cost = sharding.allgather_rows(mine["cost"], total)
    iters = sharding.allgather_rows(mine["iterations"], total)
    # device-timing convention of bench.py: the step time is the max over ranks
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        np.savez(out_path, X=X, U=U, cost=cost, iters=iters, tmax=t.numpy(), total=sharding.ordered_total(cost))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_solve_matches_single_process(tmp_path):
    from conftest import HostEmulation, random_x0

    total, world = 37, 2  # odd on purpose: shards of 19 and 18
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(world, _free_port(), total, out), nprocs=world, join=True)
    g = np.load(out)
    x0 = random_x0(2, total, seed=123)
    ref = HostEmulation().solve(2, x0, np.zeros((total, 10, 4)), 100, 1e-5)
    assert np.array_equal(g["X"], ref["X"]) and np.array_equal(g["U"], ref["U"])
    assert np.array_equal(g["cost"], ref["cost"]) and np.array_equal(g["iters"], ref["iterations"])
    assert g["tmax"][0] == 2.0
    tot = 0.0
    for c in ref["cost"]:
        tot += float(c)
    assert float(g["total"]) == tot


def test_shard_bounds_cover_everything():
    from multi_agent_solver_b200 import sharding

    for total in (0, 1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            edges = [sharding.shard_bounds(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1 and sizes == sharding.shard_counts(total, world)
