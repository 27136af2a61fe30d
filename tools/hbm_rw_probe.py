"""HBM bandwidth by direction on this GPU (torch kernels): write-only (fill), read-only (sum), copy (read + write). The roofline
denominator MEASURED_PEAKS.json uses is the copy figure; write-dominated layers (MobileNetV2's 1x1 expansions) see the fill one."""
import torch
x = torch.empty(1 << 30, dtype=torch.float16, device="cuda")
y = torch.empty_like(x)
def t(f, n=10):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
gb = x.numel() * 2 / 1e9
print("fill  (write only)  %.0f GB/s" % (gb / (t(lambda: x.fill_(1.0)) * 1e-3)))
print("sum   (read only)   %.0f GB/s" % (gb / (t(lambda: x.sum()) * 1e-3)))
print("copy  (read+write)  %.0f GB/s" % (2 * gb / (t(lambda: y.copy_(x)) * 1e-3)))
