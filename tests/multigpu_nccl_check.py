"""Multi-GPU check, run under torchrun on a box with >= 2 GPUs (tests/test_gpu_multigpu.py launches it):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_nccl_check.py

The library's own NCCL communicator (mas_b200_context_init_nccl) all-gathers every agent's (X, U, cost) after every
outer round of a Nash strategy (mas_b200_strategy_run), so that every rank holds the joint trajectory set:
  1. scenarios sharded (BASELINE configs[1] / [3] shapes): per-rank results gathered with torch.distributed ==
     the same scenarios solved unsharded on rank 0, bit for bit; the library's joint set
     (mas_b200_strategy_get_joint) == that gathered set on EVERY rank;
  2. agents sharded (configs[3]: one scenario, world x 16 LQR agents with distinct initial states): the joint
     total_cost every rank reports == the unsharded scenario's total (block-order sum), bit for bit;
  3. ranks that disagree on the shape get MAS_B200_ERR_INVALID_ARGUMENT on every rank instead of a hang.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import multi_agent_solver_b200 as mas  # noqa: E402
from multi_agent_solver_b200 import sharding  # noqa: E402


def gather_np(a, world):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    return np.stack([p.cpu().numpy() for p in parts], 0)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = mas.Context(local)
    # the library's own communicator: unique id from rank 0, distributed with torch.distributed
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.frombuffer(bytearray(mas.Context.nccl_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(uid, 0)
    ctx.init_nccl(bytes(uid.cpu().numpy().tobytes()), rank, world)
    plain = mas.Context(local)  # no communicator: unsharded runs

    ok = True

    def report(what, same):
        nonlocal ok
        ok = ok and bool(same)
        if rank == 0 or not same:
            print(f"[rank {rank}] {what}: {bool(same)}", flush=True)

    prm = mas.IlqrParams.make(100, 1e-5)
    # ---- 1. scenarios sharded ---------------------------------------------------------------------------------
    for strategy, model, S, A, outer in ((mas.Strategy.SEQUENTIAL, mas.Model.LQR4, 8 * world, 16, 10),
                                         (mas.Strategy.TRUSTREGION, mas.Model.SINGLE_TRACK_CIRC, 6 * world, 3, 6),
                                         (mas.Strategy.LINESEARCH, mas.Model.LQR4, 4 * world, 5, 4)):
        rng = np.random.default_rng(99)
        if model == mas.Model.LQR4:
            x0 = rng.uniform(-1, 1, (S, A, 4))
        else:
            th = 2.0 * np.pi * np.arange(A) / A + rng.uniform(0, 0.3, (S, 1))
            x0 = np.stack([20 * np.cos(th), 20 * np.sin(th), 1.57 + th, np.full((S, A), 4.0)], -1)
        desc = mas.example_desc(model)
        lo, hi = sharding.shard_bounds(S, rank, world)
        mine = mas.strategy_run(ctx, strategy, desc, prm, outer, x0[lo:hi])
        joint = ctx.strategy_joint(world, hi - lo, A, desc)
        st = ctx.exchange_stats()
        full = mas.strategy_run(plain, strategy, desc, prm, outer, x0)
        for k in ("X", "U", "costs", "total_cost"):
            g = gather_np(mine[k], world)
            report(f"strategy {strategy} {k}: sharded == unsharded", np.array_equal(g.reshape(full[k].shape), full[k]))
        for k, fk in (("X", "X"), ("U", "U"), ("costs", "costs")):
            report(f"strategy {strategy} joint {k} on rank {rank} == all ranks' results", np.array_equal(joint[k].reshape(full[fk].shape), full[fk]))
        if rank == 0:
            print(f"   exchange: {st['rounds']} rounds, {st['collective_ms']:.3f} ms in collectives, {st['bytes_per_round']} B received per round", flush=True)

    # ---- 2. agents sharded: one scenario, world x 16 agents ----------------------------------------------------
    A_loc, outer = 16, 10
    rng = np.random.default_rng(5)
    x0 = rng.uniform(-1, 1, (1, world * A_loc, 4))
    desc = mas.example_desc(mas.Model.LQR4)
    ctx.set_agent_sharding(True)
    mine = mas.strategy_run(ctx, mas.Strategy.SEQUENTIAL, desc, prm, outer, x0[:, rank * A_loc:(rank + 1) * A_loc])
    ctx.set_agent_sharding(False)
    full = mas.strategy_run(plain, mas.Strategy.SEQUENTIAL, desc, prm, outer, x0)
    report("agents sharded: joint total_cost == unsharded total_cost (bit-exact)", mine["total_cost"][0] == full["total_cost"][0])
    report("agents sharded: local costs == the unsharded scenario's block", np.array_equal(mine["costs"][0], full["costs"][0, rank * A_loc:(rank + 1) * A_loc]))

    # ---- 3. shape mismatch across ranks is an error everywhere, not a hang --------------------------------------
    try:
        mas.strategy_run(ctx, mas.Strategy.SEQUENTIAL, desc, prm, 3 if rank == 0 else 4, x0[:, :A_loc])
        report("mismatched max_outer rejected", False)
    except mas.MasB200Error as e:
        report("mismatched max_outer rejected", e.code != 0)

    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTIGPU NCCL CHECK", "OK" if int(flag.item()) == 1 else "FAILED", flush=True)
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
