"""Micro-measurement: copy bandwidth for L2-resident (12 MB) vs HBM-sized (1 GB) fp32 arrays (torch copy_)."""
import torch
for mb in (12, 36, 96, 1024):
    n = mb * 1024 * 1024 // 4
    a = torch.empty(n, device="cuda"); b = torch.empty(n, device="cuda")
    for _ in range(5): b.copy_(a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 200 if mb < 200 else 20
    e0.record()
    for _ in range(reps): b.copy_(a)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / reps * 1e-3
    print("%5d MB copy: %.2f us  %.0f GB/s (read+write)" % (mb, t * 1e6, 2 * n * 4 / t / 1e9))
