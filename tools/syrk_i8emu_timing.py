"""Round-2 tool: time the EXPERIMENTAL int8-emulated syrk (csrc/npw_ozaki_i8.cu) against the native DMMA kernel on one
4096^3 tile update, and report accuracy.  Run under `timeout` (the kernel has never executed):
    NPW_B200_EXPERIMENTAL=1 timeout 120 python tools/syrk_i8emu_timing.py [digits=6] [size=4096]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from numpywren_b200 import kernels  # noqa: E402


def timed(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record(); torch.cuda.synchronize()
    return out, a.elapsed_time(b) / reps


def main():
    digits = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(0)
    x = torch.randn(n, n, generator=g, dtype=torch.float64).to(dev)
    y = torch.randn(n, n, generator=g, dtype=torch.float64).to(dev)
    s = torch.randn(n, n, generator=g, dtype=torch.float64).to(dev)
    (xd, xe), t_split = timed(lambda: kernels.split_i8(x, digits))
    yd, ye = kernels.split_i8(y, digits)
    c, t_emu = timed(lambda: kernels.syrk_i8emu(s, xd, xe, yd, ye))
    ref, t_nat = timed(lambda: kernels.syrk(s, x, y))
    err = float((c - ref).norm() / ref.norm())
    pairs = digits * (digits + 1) // 2
    print(json.dumps({"size": n, "digits": digits, "int8_products": pairs, "split_ms_per_operand": t_split, "syrk_i8emu_ms": t_emu,
                      "native_syrk_ms": t_nat, "speedup_excluding_split": t_nat / t_emu, "rel_err_vs_native": err,
                      "int8_tops": 2.0 * n ** 3 * pairs / (t_emu * 1e-3) / 1e12,
                      "fp64_equivalent_tflops": 2.0 * n ** 3 / (t_emu * 1e-3) / 1e12}))


if __name__ == "__main__":
    main()
