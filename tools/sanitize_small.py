"""Small-shape invocations of every kernel for compute-sanitizer (memcheck / racecheck) runs."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numpywren_b200 import kernels  # noqa: E402

dev = torch.device("cuda:0")
rs = np.random.RandomState(0)


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


for (m, n, k) in [(128, 128, 64), (200, 136, 72), (33, 17, 5), (256, 256, 256)]:
    s, x, y = rs.randn(m, n), rs.randn(m, k), rs.randn(n, k)
    got = kernels.syrk(t(s), t(x), t(y))
    assert np.allclose(got.cpu().numpy(), s - x @ y.T)
    assert np.allclose(kernels.gemm(t(x), t(y.T.copy())).cpu().numpy(), x @ y.T)
for n in (8, 100, 128, 200, 384):
    x = rs.randn(n, n + 4)
    a = x @ x.T + n * np.eye(n)
    L, info, inv = kernels.chol_async(t(a))
    assert np.allclose(L.cpu().numpy(), np.linalg.cholesky(a))
    b = rs.randn(40, n)
    X = kernels.trsm_with_inverse(L, t(b), inv)
    assert np.allclose(X.cpu().numpy() @ np.linalg.cholesky(a).T, b)
    X2 = kernels.trsm(L, t(b))
    assert np.allclose(X2.cpu().numpy(), X.cpu().numpy())
# incl. a 16-CTA panel grid (two-level gather with two groups), several panels (TN product, rank update, T blocks) and an
# unaligned leading dimension (generic TN kernel, scalar rank-update path)
for (m, n) in [(64, 32), (300, 70), (1024, 64), (2048, 96), (600, 160), (130, 33)]:
    a = rs.randn(m, n)
    V, T, R = kernels.qr_factor(t(a))
    assert np.allclose(np.abs(R.cpu().numpy()), np.abs(np.linalg.qr(a)[1]))
# fp64 syrk emulated on the int8 tensor cores (tcgen05 / TMEM / TMA 3-D)
x, y, sm = rs.randn(256, 512), rs.randn(192, 512), rs.randn(256, 192)
xd, xe = kernels.split_i8(t(x), 7)
yd, ye = kernels.split_i8(t(y), 7)
c = kernels.syrk_i8emu(t(sm), xd, xe, yd, ye).cpu().numpy()
assert np.linalg.norm(c - (sm - x @ y.T)) / np.linalg.norm(sm - x @ y.T) < 1e-11
p = [t(rs.randn(50, 30)) for _ in range(4)]
kernels.add_matrices(*p); kernels.mul(p[0], p[1]); kernels.transpose(p[0])
torch.cuda.synchronize()
print("sanitize_small ok")
