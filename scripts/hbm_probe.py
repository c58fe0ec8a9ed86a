"""HBM ceilings for the access patterns of this path (diagnostic): pure write (fill), copy, read (sum)."""
import torch
dev = torch.device("cuda:0")
n = 3 * 1024**3  # floats: 12 GiB
x = torch.empty(n, dtype=torch.float32, device=dev)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best
ms = t(lambda: x.fill_(1.0)); print(f"fill_ 12 GiB: {ms:.3f} ms  {4*n/ms/1e6:.0f} GB/s write")
ms = t(lambda: x.zero_()); print(f"zero_ 12 GiB: {ms:.3f} ms  {4*n/ms/1e6:.0f} GB/s write")
y = torch.empty(n // 2, dtype=torch.float32, device=dev)
ms = t(lambda: y.copy_(x[: n // 2])); print(f"copy 6 GiB: {ms:.3f} ms  {4*n/ms/1e6:.0f} GB/s r+w")
ms = t(lambda: x.sum()); print(f"sum 12 GiB: {ms:.3f} ms  {4*n/ms/1e6:.0f} GB/s read")
