"""Developer GPU probe: BASELINE configs 4 (TSQR 4194304x512, tile 65536x512) and 5 (GEMM tile 8192) on ONE GPU
(config 5 scaled to N=32768 so that A, B and the Temp tree fit in one B200's HBM)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numpywren_b200 import job_runner, kernels  # noqa: E402
from numpywren_b200 import lambdapack as lp  # noqa: E402
from numpywren_b200.alg_wrappers import gemm, tsqr  # noqa: E402
from numpywren_b200.matrix import BigMatrix  # noqa: E402

dev = torch.device("cuda:0")


def run(program, **kw):
    program.start()
    torch.cuda.synchronize()
    t0 = time.time()
    job_runner.lambdapack_run(program, timeout=600, **kw)
    torch.cuda.synchronize()
    assert program.program_status() == lp.PS.SUCCESS
    return time.time() - t0


def tsqr_case(m, n, b):
    nb = m // b
    X = BigMatrix(f"t4_{m}", shape=(m, n), shard_sizes=(b, n))
    X.free()
    for j in range(nb):
        t = torch.empty(b, n, dtype=torch.float64, device=dev)
        kernels.fill_random(t, seed=3, row0=j * b)
        X._put_block_ref(t, j, 0)
    for rep in range(2):
        program, meta = tsqr(X)
        dt = run(program)
        flops = 2.0 * m * n * n - 2.0 * n ** 3 / 3.0
        Rs = meta["outputs"][0]
        levels = int(np.ceil(np.log2(nb)))
        R = Rs.get_block(levels, 0)
        print(f"TSQR {m}x{n} tile ({b},{n}): {len(program.program.nodes)} nodes, {dt * 1e3:.1f} ms -> {flops / dt * 1e-12:.2f} TFLOP/s,"
              f" {2 * m * n * 8 / dt * 1e-9:.0f} GB/s of A+V")
        if rep == 1:
            # check R^T R = X^T X (column-norm-wise) on the device
            G = torch.zeros(n, n, dtype=torch.float64, device=dev)
            for j in range(nb):
                t = X._get_block_ref(j, 0)
                G += t.T @ t
            err = float((R.T @ R - G).norm() / G.norm())
            print(f"   ||R^T R - X^T X|| / ||X^T X|| = {err:.2e}")
        for mm in meta["outputs"]:
            mm.free()
    X.free()


def gemm_case(n, b):
    nb = n // b
    A = BigMatrix(f"g5a_{n}", shape=(n, n), shard_sizes=(b, b)); A.free()
    B = BigMatrix(f"g5b_{n}", shape=(n, n), shard_sizes=(b, b)); B.free()
    for i in range(nb):
        for k in range(nb):
            ta = torch.empty(b, b, dtype=torch.float64, device=dev); kernels.fill_random(ta, 1, i * b, k * b)
            tb = torch.empty(b, b, dtype=torch.float64, device=dev); kernels.fill_random(tb, 2, i * b, k * b)
            A._put_block_ref(ta, i, k); B._put_block_ref(tb, i, k)
    for rep in range(2):
        program, meta = gemm(A, B)
        dt = run(program)
        print(f"GEMM N={n} tile {b}: {len(program.program.nodes)} nodes, {dt * 1e3:.1f} ms -> {2.0 * n ** 3 / dt * 1e-12:.2f} TFLOP/s")
        if rep == 1:
            C00 = meta["outputs"][0].get_block(0, 0)
            ref = sum(A._get_block_ref(0, k) @ B._get_block_ref(k, 0) for k in range(nb))
            print(f"   C[0,0] rel err vs cuBLAS {float((C00 - ref).norm() / ref.norm()):.2e}")
        for mm in meta["outputs"] + meta["intermediates"]:
            mm.free()


if __name__ == "__main__":
    tsqr_case(262144, 512, 65536)
    tsqr_case(4194304, 512, 65536)
    gemm_case(16384, 8192)
    gemm_case(32768, 8192)
