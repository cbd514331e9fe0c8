"""Developer GPU probe (run under gpurun): kernel correctness vs torch/cuBLAS fp64 and raw timings.

Not part of the product or the test-suite; writes gpurun_out/gpu_check.json.
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numpywren_b200 import _capi, kernels  # noqa: E402

OUT = {}


def rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts) / len(ts)


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    print(torch.cuda.get_device_name(0), "host cpus", os.cpu_count())
    OUT["device"] = torch.cuda.get_device_name(0)
    OUT["cpus"] = os.cpu_count()
    try:
        import psutil
        OUT["host_mem_gb"] = psutil.virtual_memory().total / 2**30
        print("host mem GB", OUT["host_mem_gb"])
    except Exception as e:  # pragma: no cover
        print("psutil", e)

    # ---- syrk / gemm NT correctness at assorted shapes (TMA path and generic path)
    errs = {}
    for (m, n, k) in [(128, 128, 16), (128, 128, 128), (256, 384, 200), (100, 70, 36), (64, 64, 64), (33, 17, 5),
                      (512, 512, 512), (1024, 1024, 1024), (130, 258, 1000)]:
        s = torch.randn(m, n, dtype=torch.float64, device=dev)
        x = torch.randn(m, k, dtype=torch.float64, device=dev)
        y = torch.randn(n, k, dtype=torch.float64, device=dev)
        ref = s - x @ y.T
        got = kernels.syrk(s, x, y)
        errs[f"syrk_{m}x{n}x{k}"] = rel(got, ref)
    for (m, n, k) in [(64, 64, 64), (200, 136, 72), (512, 512, 512)]:
        a = torch.randn(m, k, dtype=torch.float64, device=dev)
        b = torch.randn(k, n, dtype=torch.float64, device=dev)
        errs[f"gemm_nn_{m}x{n}x{k}"] = rel(kernels.gemm(a, b), a @ b)
        errs[f"gemm_tn_{m}x{n}x{k}"] = rel(kernels.gemm(a.T.contiguous(), b, transpose_A=True), a @ b)
        errs[f"gemm_view_{m}x{n}x{k}"] = rel(kernels.gemm(a, b.T.contiguous().T), a @ b)
    print(json.dumps(errs, indent=1))
    OUT["gemm_errs"] = errs

    # ---- chol / trsm
    for n in [8, 64, 128, 200, 256, 1024, 4096]:
        x = torch.randn(n, n + 8, dtype=torch.float64, device=dev)
        a = x @ x.T + n * torch.eye(n, dtype=torch.float64, device=dev)
        L = kernels.chol(a)
        Lref = torch.linalg.cholesky(a)
        e_chol = rel(L, Lref)
        b = torch.randn(max(n // 2, 3), n, dtype=torch.float64, device=dev)
        X = kernels.trsm(Lref, b)
        Xref = torch.linalg.solve_triangular(Lref.T, b, upper=True, left=False)
        e_trsm = rel(X, Xref)
        Lq, info, inv = kernels.chol_async(a)
        X2 = kernels.trsm_with_inverse(Lq, b, inv)
        e_trsm2 = rel(X2, Xref)
        print(f"n={n} chol {e_chol:.2e} trsm {e_trsm:.2e} trsm(inv) {e_trsm2:.2e} info {int(info.item())}")
        OUT[f"chol_{n}"] = e_chol
        OUT[f"trsm_{n}"] = e_trsm
    # non-SPD detection
    bad = -torch.eye(64, dtype=torch.float64, device=dev)
    try:
        kernels.chol(bad)
        print("non-SPD: NO ERROR (bad)")
        OUT["nonspd"] = "no error"
    except Exception as e:
        print("non-SPD raised", type(e).__name__)
        OUT["nonspd"] = type(e).__name__

    # ---- elementwise
    t = [torch.randn(300, 257, dtype=torch.float64, device=dev) for _ in range(4)]
    print("addn exact", bool(torch.equal(kernels.add_matrices(*t), (torch.zeros_like(t[0]) + t[0] + t[1] + t[2] + t[3]))))
    print("transpose exact", bool(torch.equal(kernels.transpose(t[0]), t[0].T.contiguous())))
    print("mul exact", bool(torch.equal(kernels.mul(t[0], t[1]), t[0] * t[1])))

    # ---- timings at the benchmark tile (4096)
    b = 4096
    s = torch.randn(b, b, dtype=torch.float64, device=dev)
    x = torch.randn(b, b, dtype=torch.float64, device=dev)
    y = torch.randn(b, b, dtype=torch.float64, device=dev)
    flops = 2.0 * b ** 3
    tmin, tavg = timeit(lambda: kernels.syrk(s, x, y), reps=10)
    print(f"npw syrk 4096: min {tmin:.3f} ms avg {tavg:.3f} ms -> {flops / tmin * 1e-9:.2f} / {flops / tavg * 1e-9:.2f} TFLOP/s")
    OUT["syrk4096_ms"] = [tmin, tavg]
    tmin, tavg = timeit(lambda: torch.addmm(s, x, y.T, beta=1.0, alpha=-1.0), reps=10)
    print(f"cuBLAS addmm 4096: min {tmin:.3f} ms avg {tavg:.3f} -> {flops / tmin * 1e-9:.2f} / {flops / tavg * 1e-9:.2f} TFLOP/s")
    OUT["cublas4096_ms"] = [tmin, tavg]
    b8 = 8192
    a8 = torch.randn(b8, b8, dtype=torch.float64, device=dev)
    c8 = torch.randn(b8, b8, dtype=torch.float64, device=dev)
    tmin, tavg = timeit(lambda: torch.matmul(a8, c8), reps=10)
    f8 = 2.0 * b8 ** 3
    print(f"cuBLAS dgemm 8192: min {tmin:.3f} ms avg {tavg:.3f} -> {f8 / tmin * 1e-9:.2f} / {f8 / tavg * 1e-9:.2f} TFLOP/s")
    OUT["cublas8192_ms"] = [tmin, tavg]
    s8 = torch.randn(b8, b8, dtype=torch.float64, device=dev)
    tmin, tavg = timeit(lambda: kernels.syrk(s8, a8, c8), reps=5)
    print(f"npw syrk 8192: min {tmin:.3f} ms avg {tavg:.3f} -> {f8 / tmin * 1e-9:.2f} / {f8 / tavg * 1e-9:.2f} TFLOP/s")
    OUT["syrk8192_ms"] = [tmin, tavg]
    # sustained 4 s cuBLAS fp64
    t0 = time.time()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); nrep = 0
    while time.time() - t0 < 4.0:
        for _ in range(4):
            torch.matmul(a8, c8); nrep += 1
        torch.cuda.synchronize()
    e1.record(); e1.synchronize()
    sus = f8 * nrep / e0.elapsed_time(e1) * 1e-9
    print(f"cuBLAS dgemm 8192 sustained: {sus:.2f} TFLOP/s over {nrep} reps")
    OUT["cublas8192_sustained_tflops"] = sus
    del a8, c8, s8

    x = torch.randn(b, b + 8, dtype=torch.float64, device=dev)
    a = x @ x.T + b * torch.eye(b, dtype=torch.float64, device=dev)
    tmin, tavg = timeit(lambda: kernels.chol_async(a), reps=5)
    print(f"npw potrf 4096: min {tmin:.3f} ms ({b**3 / 3 / tmin * 1e-9:.2f} TFLOP/s)")
    OUT["potrf4096_ms"] = [tmin, tavg]
    tmin, tavg = timeit(lambda: torch.linalg.cholesky(a), reps=5)
    print(f"cusolver potrf 4096: min {tmin:.3f} ms")
    L, info, inv = kernels.chol_async(a)
    tmin, tavg = timeit(lambda: kernels.trsm_with_inverse(L, s, inv), reps=5)
    print(f"npw trsm 4096: min {tmin:.3f} ms ({b**3 / tmin * 1e-9:.2f} TFLOP/s)")
    OUT["trsm4096_ms"] = [tmin, tavg]
    tmin, tavg = timeit(lambda: torch.linalg.solve_triangular(L.T, s, upper=True, left=False), reps=5)
    print(f"cuBLAS trsm 4096: min {tmin:.3f} ms")

    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/gpu_check.json", "w") as f:
        json.dump(OUT, f, indent=1)


if __name__ == "__main__":
    main()
