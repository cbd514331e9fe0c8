"""Launch the dominant kernel (kernels.syrk on 4096^2 fp64 tiles) a few dozen times: the ncu capture target."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numpywren_b200 import kernels  # noqa: E402

b = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
dev = torch.device("cuda:0")
sets = []
for i in range(4):
    t = [torch.empty(b, b, dtype=torch.float64, device=dev) for _ in range(3)]
    for j, x in enumerate(t):
        kernels.fill_random(x, 100 + 10 * i + j)
    sets.append(t)
for r in range(reps):
    s, x, y = sets[r % 4]
    kernels.syrk(s, x, y, out=s)
torch.cuda.synchronize()
print("done", reps)
