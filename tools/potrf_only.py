"""ncu target: kernels.chol on one tile (default 512) so that potf2_inv_kernel launches can be captured."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numpywren_b200 import kernels  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
x = torch.randn(n, n + 8, dtype=torch.float64, device="cuda:0")
a = x @ x.T + n * torch.eye(n, dtype=torch.float64, device="cuda:0")
for _ in range(3):
    kernels.chol_async(a)
torch.cuda.synchronize()
print("done")
