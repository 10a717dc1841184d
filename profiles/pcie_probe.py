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


# ---- the e2e step's shapes: 23.5 MB up (U), 7.8 MB up (p) while 7.8 MB goes down ----
def timed(fn, reps=8):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


nU, nP = 23_518_296, 7_839_432
hU = torch.empty(nU, dtype=torch.uint8).pin_memory()
dU = torch.empty(nU, dtype=torch.uint8, device='cuda')
hP = torch.empty(nP, dtype=torch.uint8).pin_memory()
dP = torch.empty(nP, dtype=torch.uint8, device='cuda')
hO = torch.empty(nP, dtype=torch.uint8).pin_memory()
dO = torch.empty(nP, dtype=torch.uint8, device='cuda')
s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()


def up_one():
    with torch.cuda.stream(s1):
        dU.copy_(hU, non_blocking=True)


def up_two():
    h = nU // 2
    with torch.cuda.stream(s1):
        dU[:h].copy_(hU[:h], non_blocking=True)
    with torch.cuda.stream(s2):
        dU[h:].copy_(hU[h:], non_blocking=True)


def p_then_out():
    with torch.cuda.stream(s1):
        dP.copy_(hP, non_blocking=True)
        hO.copy_(dO, non_blocking=True)


def p_and_out():
    with torch.cuda.stream(s1):
        dP.copy_(hP, non_blocking=True)
    with torch.cuda.stream(s3):
        hO.copy_(dO, non_blocking=True)


print('U 23.5 MB up, one stream: %.3f ms; two streams: %.3f ms' % (timed(up_one) * 1e3, timed(up_two) * 1e3))
src = torch.ones(nU, dtype=torch.uint8)


def hot(fn):
    def g():
        hU.copy_(src)        # the CPU has just written the source (the solver's situation); not timed separately
        fn()
    return g


t_fill = timed(lambda: hU.copy_(src))
print('  source just rewritten by the CPU: one stream %.3f ms, two streams %.3f ms (incl. %.3f ms refill)'
      % (timed(hot(up_one)) * 1e3, timed(hot(up_two)) * 1e3, t_fill * 1e3))
print('p 7.8 MB up then 7.8 MB down, serial: %.3f ms; concurrent (full duplex): %.3f ms' % (timed(p_then_out) * 1e3, timed(p_and_out) * 1e3))
