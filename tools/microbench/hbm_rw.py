"""HBM bandwidth by access mix on this GPU: write-only (fill), read-only (sum), copy (read+write)."""
import torch
n = 1 << 30   # 4 GiB of int32
a = torch.empty(n, dtype=torch.int32, device="cuda")
b = torch.empty(n, dtype=torch.int32, device="cuda")
def t(f, reps=5):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: a.fill_(7)); print("write-only  %.0f GB/s" % (4 * n / ms / 1e6))
ms = t(lambda: a.sum()); print("read-only   %.0f GB/s" % (4 * n / ms / 1e6))
ms = t(lambda: b.copy_(a)); print("copy r+w    %.0f GB/s (total bytes)" % (8 * n / ms / 1e6))
