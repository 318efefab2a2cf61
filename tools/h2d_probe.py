"""Development aid: host->device copy rate of the pinned 512^3 float volume, alone and split over two streams."""
import time
import torch
n = 512 ** 3
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
one = t(lambda: d.copy_(h, non_blocking=True))
def two():
    with torch.cuda.stream(s1):
        d[: n // 2].copy_(h[: n // 2], non_blocking=True)
    with torch.cuda.stream(s2):
        d[n // 2:].copy_(h[n // 2:], non_blocking=True)
def chunks16():
    c = n // 16
    for i in range(16):
        d[i * c:(i + 1) * c].copy_(h[i * c:(i + 1) * c], non_blocking=True)
print(f"one copy: {one*1e3:.2f} ms = {n*4/one/1e9:.1f} GB/s; two streams: {t(two)*1e3:.2f} ms; 16 chunks: {t(chunks16)*1e3:.2f} ms")
back = torch.empty(n, dtype=torch.float32).pin_memory()
print(f"d2h: {t(lambda: back.copy_(d, non_blocking=True))*1e3:.2f} ms")
