"""Launch kernels.qr_factor on one TSQR leaf (65536 x 512): ncu target + timing."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numpywren_b200 import kernels  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
a = torch.empty(m, n, dtype=torch.float64, device="cuda:0")
kernels.fill_random(a, 5)
for r in range(reps):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    V, T, R = kernels.qr_factor(a)
    e1.record(); e1.synchronize()
    print(f"qr_factor {m}x{n}: {e0.elapsed_time(e1):.2f} ms -> {(2.0 * m * n * n - 2.0 * n ** 3 / 3) / e0.elapsed_time(e1) * 1e-9:.2f} TFLOP/s")
