"""Developer GPU probe: potrf / trsm tile timings and accuracy after kernel changes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numpywren_b200 import kernels  # noqa: E402


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


dev = torch.device("cuda:0")
for n in (128, 1024, 4096):
    x = torch.randn(n, n + 8, dtype=torch.float64, device=dev)
    a = x @ x.T + n * torch.eye(n, dtype=torch.float64, device=dev)
    L, info, inv = kernels.chol_async(a)
    ref = torch.linalg.cholesky(a)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    X = kernels.trsm_with_inverse(L, b, inv)
    Xr = torch.linalg.solve_triangular(ref.T, b, upper=True, left=False)
    print(f"n={n}: chol err {float((L - ref).norm() / ref.norm()):.2e} trsm err {float((X - Xr).norm() / Xr.norm()):.2e} info {int(info.item())}"
          f" | potrf {timeit(lambda: kernels.chol_async(a)):.3f} ms trsm {timeit(lambda: kernels.trsm_with_inverse(L, b, inv)):.3f} ms"
          f" trsm(no inv) {timeit(lambda: kernels.trsm_with_inverse(L, b, None)):.3f} ms"
          f" | cusolver potrf {timeit(lambda: torch.linalg.cholesky(a)):.3f} ms")
