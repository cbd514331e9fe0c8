"""Developer GPU probe: phase breakdown of qr_panel_reg_kernel (library built with -DNPW_QR_PROFILE, see
tools/r02_call25.sh).  Calls npw_geqrt_f64 directly with its own work buffer and reads the per-CTA phase counters that the
profile build accumulates behind the packet area."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from numpywren_b200 import _capi, kernels  # noqa: E402

PK_BYTES = 179200
lib = _capi.load()
m, n = 65536, 512
a = torch.empty(m, n, dtype=torch.float64, device="cuda:0")
kernels.fill_random(a, 5)
T = torch.empty(n, n, dtype=torch.float64, device="cuda:0")
R = torch.empty(n, n, dtype=torch.float64, device="cuda:0")
work = torch.zeros(lib.npw_geqrt_work_bytes(m, n) // 8 + 1, dtype=torch.float64, device="cuda:0")
for rep in range(2):
    v = a.clone()
    rc = lib.npw_geqrt_f64(v.data_ptr(), n, T.data_ptr(), n, R.data_ptr(), n, v.data_ptr(), n, m, n, work.data_ptr(),
                           torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc
    torch.cuda.synchronize()
raw = work.view(torch.int64)[(PK_BYTES + 256) // 8:(PK_BYTES + 256) // 8 + 160 * 2 * 8].cpu().numpy().reshape(160, 2, 8)[:148]
names = ["leader stage", "poll groups", "barrier+sum", "scalars", "pass", "blocksum+publish", "-", "other (publish_row etc.)"]
steps = 16 * 33
for who, sel in (("leaders (cta % 12 == 0)", [c for c in range(148) if c % 12 == 0]), ("others", [c for c in range(148) if c % 12]),
                 ("cta 0", [0])):
    for th, tn in ((0, "thread 0"), (1, "thread 255")):
        x = raw[sel, th].mean(axis=0) / steps
        print(f"{who:26s} {tn:10s} cycles per step: " + ", ".join(f"{nm} {v:.0f}" for nm, v in zip(names, x) if nm != "-") + f" | total {x.sum():.0f}")
