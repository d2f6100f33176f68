"""D2H copy bandwidth while the SMs are busy with (a) nothing, (b) an fp64 GEMM loop, (c) a streaming add loop."""
import threading
import time

import torch

n = 170 * 1024 * 1024
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
host = torch.empty(n, dtype=torch.uint8).pin_memory()
cs = torch.cuda.Stream()
ws = torch.cuda.Stream()
a = torch.randn(4096, 4096, dtype=torch.float64, device="cuda")
big = torch.zeros(1 << 28, dtype=torch.float64, device="cuda")  # 2 GB


def copy_rate(seconds=0.5):
    cnt = 0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(cs):
        while time.perf_counter() - t0 < seconds:
            host.copy_(dev, non_blocking=True)
            cs.synchronize()
            cnt += 1
    return cnt * n / (time.perf_counter() - t0) / 1e9


def load(kind, stop, counter):
    with torch.cuda.stream(ws):
        while not stop.is_set():
            for _ in range(4):
                if kind == "gemm":
                    torch.mm(a, a)
                else:
                    big.add_(1.0)
            ws.synchronize()
            counter[0] += 4


print("idle GPU: D2H %.1f GB/s" % copy_rate())
for kind in ("gemm", "stream"):
    stop, counter = threading.Event(), [0]
    th = threading.Thread(target=load, args=(kind, stop, counter))
    th.start()
    time.sleep(0.3)
    c0, t0 = counter[0], time.perf_counter()
    r = copy_rate(1.0)
    rate_with = (counter[0] - c0) / (time.perf_counter() - t0)
    c0, t0 = counter[0], time.perf_counter()
    time.sleep(1.0)
    rate_without = (counter[0] - c0) / (time.perf_counter() - t0)
    stop.set()
    th.join()
    print(f"{kind}: D2H {r:.1f} GB/s; load ops/s with copy {rate_with:.1f}, without {rate_without:.1f}")
