"""Pure-write HBM bandwidth (fill of a 4 GB buffer) next to the copy bandwidth: the denominator that matters for a
store-only kernel such as path generation."""
import torch
x = torch.empty(1 << 30, dtype=torch.float32, device="cuda")
y = torch.empty_like(x)
def timed(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = timed(lambda: x.fill_(1.0))
print(f"fill  4 GiB: {ms:.3f} ms  {x.numel()*4/ms/1e6:.0f} GB/s written")
ms = timed(lambda: y.copy_(x))
print(f"copy  4 GiB: {ms:.3f} ms  {2*x.numel()*4/ms/1e6:.0f} GB/s read+written")
ms = timed(lambda: x.sum())
print(f"read  4 GiB: {ms:.3f} ms  {x.numel()*4/ms/1e6:.0f} GB/s read")
