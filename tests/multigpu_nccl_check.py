"""Multi-GPU check, run under torchrun on a box with >= 2 GPUs (not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_nccl_check.py

Every rank owns a contiguous shard of scenarios of a sequential-Nash LQR run (BASELINE configs[3]) and a
trust-region run (configs[1]); the library's NCCL communicator all-gathers (X, U, cost) after every outer round
(mas_b200_context_init_nccl + mas_b200_strategy_run).  Rank results are gathered with torch.distributed and
compared with the same scenarios solved unsharded on rank 0: bit-identical per scenario, whatever the sharding.
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

    ok = True
    for strategy, model, S, A, outer in ((mas.Strategy.SEQUENTIAL, mas.Model.LQR4, 8 * world, 16, 10),
                                         (mas.Strategy.TRUSTREGION, mas.Model.SINGLE_TRACK_CIRC, 6 * world, 3, 6)):
        rng = np.random.default_rng(99)
        if model == mas.Model.LQR4:
            x0 = rng.uniform(-1, 1, (S, A, 4))
        else:
            th = 2.0 * np.pi * np.arange(A) / A + rng.uniform(0, 0.3, (S, 1))
            x0 = np.stack([20 * np.cos(th), 20 * np.sin(th), 1.57 + th, np.full((S, A), 4.0)], -1)
        desc = mas.example_desc(model)
        prm = mas.IlqrParams.make(100, 1e-5)
        lo, hi = sharding.shard_bounds(S, rank, world)
        mine = mas.strategy_run(ctx, strategy, desc, prm, outer, x0[lo:hi])
        gathered = {}
        for k in ("X", "U", "costs", "total_cost"):
            t = torch.from_numpy(np.ascontiguousarray(mine[k])).cuda()
            parts = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(parts, t)
            gathered[k] = torch.cat(parts, 0).cpu().numpy()
        if rank == 0:
            plain = mas.Context(local)  # no communicator: unsharded reference run
            full = mas.strategy_run(plain, strategy, desc, prm, outer, x0)
            for k in ("X", "U", "costs", "total_cost"):
                same = np.array_equal(gathered[k], full[k])
                ok = ok and same
                print(f"strategy {strategy} {k}: sharded == unsharded: {same}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTIGPU NCCL CHECK", "OK" if ok else "FAILED", flush=True)
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
