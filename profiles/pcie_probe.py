"""PCIe probe: pinned H2D / D2H bandwidth with 1, 2, 4 concurrent copies, and a zero-copy kernel read (torch only)."""
import time
import torch

n = 39_197_160
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device='cuda')
torch.cuda.synchronize()


def bw(k, d2h=False):
    streams = [torch.cuda.Stream() for _ in range(k)]
    chunk = n // k
    best = 0.0
    for rep in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                a, b = i * chunk, (i + 1) * chunk if i < k - 1 else n
                if d2h:
                    h[a:b].copy_(d[a:b], non_blocking=True)
                else:
                    d[a:b].copy_(h[a:b], non_blocking=True)
        torch.cuda.synchronize()
        best = max(best, n / (time.perf_counter() - t0) / 1e9)
    return best


for k in (1, 2, 4):
    print('H2D %d stream(s): %.1f GB/s   D2H: %.1f GB/s' % (k, bw(k), bw(k, True)))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
big = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
dbig = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for _ in range(3):
    e0.record(); dbig.copy_(big, non_blocking=True); e1.record(); torch.cuda.synchronize()
    print('H2D 256 MiB: %.1f GB/s' % ((256 << 20) / e0.elapsed_time(e1) / 1e6))
