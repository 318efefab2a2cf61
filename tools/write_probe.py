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
# eight separate 512 MiB planes written by ONE kernel with interleaved blocks (torch._foreach_zero_): the measures
# kernel's output pattern (8 planes 537 MB apart)
outs = [torch.empty(1 << 27, dtype=torch.float32, device="cuda") for _ in range(8)]
ms = t(lambda: torch._foreach_zero_(outs)); print(f"8 planes of 512 MiB zeroed by one multi-tensor kernel: {ms:.3f} ms = {8*(1<<29)/ms/1e6:.0f} GB/s written")
big = torch.empty(8, 1 << 27, dtype=torch.float32, device="cuda")
ms = t(lambda: big.transpose(0, 1).zero_() if False else big.zero_()); print(f"same 4 GiB as one tensor: {ms:.3f} ms = {8*(1<<29)/ms/1e6:.0f} GB/s")
src = torch.empty(1 << 27, dtype=torch.float32, device="cuda")
def nine():
    # one plane read, eight planes written, each element's 8 outputs issued together (the measures kernel's pattern)
    torch.stack([src] * 8, out=big)
ms = t(nine); print(f"1 plane read + 8 planes written (torch.stack): {ms:.3f} ms = {(9*(1<<29))/ms/1e6:.0f} GB/s moved")
