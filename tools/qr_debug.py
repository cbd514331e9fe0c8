"""Developer GPU probe: kernels.qr_factor vs the oracle on a few shapes, component-wise errors (V, T, R)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from numpywren_b200 import kernels  # noqa: E402
from oracle import npw_oracle as orc  # noqa: E402  (checker only)


def rel(x, y):
    x = x.cpu().numpy()
    return float(np.linalg.norm(x - y) / max(np.linalg.norm(y), 1e-300))


for m, n in [(16, 16), (64, 8), (96, 32), (200, 40), (512, 64), (4096, 64), (2048, 512), (20000, 96)]:
    a = np.random.RandomState(m + n).randn(m, n)
    V, T, R = kernels.qr_factor(torch.from_numpy(a).to("cuda:0"))
    torch.cuda.synchronize()
    v, t, r = orc.qr_factor(a)
    print(f"{m}x{n}: R {rel(R, r):.2e} V {rel(V, v):.2e} T {rel(T, t):.2e} nan {bool(torch.isnan(T).any())}", flush=True)
