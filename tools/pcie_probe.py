"""Host<->device copy bandwidth of this box through pinned memory (what bounds bench.py's e2e leg).

    python tools/pcie_probe.py                                            # one GPU
    python -m torch.distributed.run --nproc-per-node 8 tools/pcie_probe.py   # all GPUs copying at the same time
"""
import json
import os

import torch


def probe(mb: int = 256, reps: int = 8, barrier=None) -> dict:
    n = mb * 1024 * 1024
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    out = {}
    for name, dst, src in (("d2h", host, dev), ("h2d", dev, host)):
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        if barrier:
            barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        out[name + "_gbps"] = n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
    return out


if __name__ == "__main__":
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        print(json.dumps(probe()))
    else:
        import torch.distributed as dist

        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
        r = probe(barrier=dist.barrier)
        t = torch.tensor([r["d2h_gbps"], r["h2d_gbps"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        if dist.get_rank() == 0:
            print(json.dumps({"gpus": world, "aggregate_d2h_gbps": float(t[0]), "aggregate_h2d_gbps": float(t[1]), "rank0": r}))
        dist.destroy_process_group()
