"""Host<->device copy bandwidth of this box through pinned memory (what bounds bench.py's e2e leg)."""
import json

import torch


def probe(mb: int = 256, reps: int = 5) -> dict:
    n = mb * 1024 * 1024
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    out = {}
    for name, dst, src in (("d2h", host, dev), ("h2d", dev, host)):
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        out[name + "_gbps"] = n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
    return out


if __name__ == "__main__":
    print(json.dumps(probe()))
