"""ncu target: the streaming tile kernels (addn / mul / copy2d / transpose / fill2d / add_diag / fill_random) on
b x b fp64 tiles, each launched a few times on operands larger than L2 in total."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numpywren_b200 import kernels  # noqa: E402

b = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda:0")
t = [torch.empty(b, b, dtype=torch.float64, device=dev) for _ in range(5)]
for i, x in enumerate(t):
    kernels.fill_random(x, 7 + i)
for r in range(2):
    kernels.add_matrices(t[0], t[1], t[2], t[3])
    kernels.mul(t[0], t[1])
    kernels.identity(t[2])
    kernels.transpose(t[3])
    kernels.add_diag(t[4], 1.0)
torch.cuda.synchronize()
print("done")
