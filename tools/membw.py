#!/usr/bin/env python
"""Write-only / read-only / copy HBM bandwidth on this GPU (context for the K1 / K3 store-bound kernels:
MEASURED_PEAKS.json's hbm_gbs is a copy, i.e. half reads and half writes)."""
import torch
n = 1 << 28                       # 1 GiB of fp32
a = torch.empty(n, device="cuda"); b = torch.empty(n, device="cuda")
def t(fn, nbytes, name, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(it):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("%-12s %.0f GB/s" % (name, nbytes / best / 1e6))
t(lambda: a.zero_(), 4 * n, "write-only")
t(lambda: a.sum(), 4 * n, "read-only")
t(lambda: b.copy_(a), 8 * n, "copy")
c = torch.empty(1 << 26, device="cuda")   # 256 MB (the size of one z tensor)
t(lambda: c.zero_(), 4 * (1 << 26), "write 256MB")
