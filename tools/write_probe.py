"""Development aid: pure-write and copy bandwidth of this B200 (what a write-dominated kernel can hope for)."""
import torch
n = 1 << 30
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")
def t(fn, reps=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: a.fill_(0.0)); print(f"fill 4 GiB: {ms:.3f} ms = {n*4/ms/1e6:.0f} GB/s written")
ms = t(lambda: a.zero_()); print(f"memset 4 GiB: {ms:.3f} ms = {n*4/ms/1e6:.0f} GB/s written")
ms = t(lambda: b.copy_(a)); print(f"copy 4 GiB: {ms:.3f} ms = {2*n*4/ms/1e6:.0f} GB/s read+write")
ms = t(lambda: a.sum()); print(f"read 4 GiB (sum): {ms:.3f} ms = {n*4/ms/1e6:.0f} GB/s read")
