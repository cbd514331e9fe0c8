"""Round-2 go/no-go probe (tools only, NOT product code): how fast and how accurate is fp64 syrk emulated with int8
tensor-core products on this GPU, using the LIBRARY int8 GEMM (torch._int_mm → cuBLASLt) for the products and torch
elementwise ops for split/combine?  It bounds from below what a hand-written tcgen05 kind::i8 kernel with a fused
combine epilogue can reach, before any kernel is written.  See tools/ozaki_prototype.py for the scheme.

  python tools/ozaki_lib_probe.py [--size 4096] [--digits 6] [--device cuda]
Prints one JSON line: per-phase milliseconds, the native-kernel time for the same update, and the relative error."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def split_rows(A, s):
    """A (m x k fp64) -> (digits: list of s int8 m x k tensors, e: m exponents); widths 6, 7, 7, ... bits."""
    amax = A.abs().amax(dim=1)
    e = torch.where(amax > 0, torch.ceil(torch.log2(torch.where(amax > 0, amax, torch.ones_like(amax)))), torch.zeros_like(amax))
    r = A * torch.exp2(-e)[:, None]
    digits = []
    for p in range(s):
        r = r * (64.0 if p == 0 else 128.0)
        q = torch.round(r)
        digits.append(q.to(torch.int8))
        r = r - q
    return digits, e


def combine(S, groups, ex, ey):
    """S - diag(2^ex) (sum_d 2^-(12 + 7 d) P_d) diag(2^ey)."""
    acc = torch.zeros_like(S)
    for d, P in groups.items():
        acc.add_(P.to(torch.float64), alpha=2.0 ** -(12 + 7 * d))
    return S - acc * torch.exp2(ex)[:, None] * torch.exp2(ey)[None, :]


def products(dx, dy, s):
    groups = {}
    for d in range(s):                                # pairs with p + q > s - 1 (0-based) are dropped
        for p in range(d + 1):
            P = torch._int_mm(dx[p], dy[d - p].t())
            groups[d] = P if d not in groups else groups[d].add_(P)
    return groups


def timed(fn, dev, reps=3):
    if dev.type != "cuda":
        t = time.time(); out = fn(); return out, (time.time() - t) * 1e3
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record(); torch.cuda.synchronize()
    return out, a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--digits", type=int, default=6)
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    a = ap.parse_args()
    dev = torch.device(a.device)
    n, s = a.size, a.digits
    g = torch.Generator(device="cpu").manual_seed(0)
    X = torch.randn(n, n, generator=g, dtype=torch.float64).to(dev)
    Y = torch.randn(n, n, generator=g, dtype=torch.float64).to(dev)
    S = torch.randn(n, n, generator=g, dtype=torch.float64).to(dev)
    (dx, ex), t_split = timed(lambda: split_rows(X, s), dev)
    dy, ey = split_rows(Y, s)
    groups, t_prod = timed(lambda: products(dx, dy, s), dev)
    C, t_comb = timed(lambda: combine(S, groups, ex, ey), dev)
    ref = S - X @ Y.t()
    out = {"size": n, "digits": s, "int8_products": s * (s + 1) // 2, "split_ms_per_operand": t_split, "products_ms": t_prod,
           "combine_ms": t_comb, "rel_err_vs_fp64": float((C - ref).norm() / ref.norm()), "device": str(dev)}
    if dev.type == "cuda":
        from numpywren_b200 import kernels
        _, out["native_syrk_ms"] = timed(lambda: kernels.syrk(S, X, Y), dev)
        out["int8_tops"] = 2.0 * n ** 3 * out["int8_products"] / (t_prod * 1e-3) / 1e12
    print(json.dumps(out))


if __name__ == "__main__":
    main()
