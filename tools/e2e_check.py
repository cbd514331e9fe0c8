"""Developer GPU probe: end-to-end Cholesky through alg_wrappers + lambdapack_run vs golden/oracle, plus timings."""
import glob
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from numpywren_b200 import job_runner, kernels  # noqa: E402
from numpywren_b200.alg_wrappers import cholesky, gemm  # noqa: E402
from numpywren_b200.matrix import BigMatrix  # noqa: E402
from numpywren_b200.matrix_init import shard_matrix  # noqa: E402
from numpywren_b200 import lambdapack as lp  # noqa: E402


def run(program):
    program.start()
    job_runner.lambdapack_run(program, timeout=600)
    assert program.program_status() == lp.PS.SUCCESS, program.program_status()


def golden_cholesky():
    for f in sorted(glob.glob(os.path.join(ROOT, "tests/golden/cholesky_*.npz"))):
        g = np.load(f)
        n, b, lam = int(g["n"]), int(g["b"]), float(g["lambdav"])
        A = BigMatrix("A_" + os.path.basename(f), shape=(n, n), shard_sizes=(b, b), lambdav=lam)
        A.free()
        shard_matrix(A, g["A"])
        program, meta = cholesky(A)
        run(program)
        L = meta["outputs"][0].numpy()
        err = np.linalg.norm(L - g["L"]) / np.linalg.norm(g["L"])
        print(os.path.basename(f), "rel err vs reference", err, "nodes", len(program.program.nodes))
        assert err < 1e-10


def golden_gemm():
    for f in sorted(glob.glob(os.path.join(ROOT, "tests/golden/gemm_*.npz"))):
        g = np.load(f)
        n, b = int(g["n"]), int(g["b"])
        A = BigMatrix("GA_" + os.path.basename(f), shape=(n, n), shard_sizes=(b, b))
        B = BigMatrix("GB_" + os.path.basename(f), shape=(n, n), shard_sizes=(b, b))
        shard_matrix(A, g["A"]); shard_matrix(B, g["B"])
        program, meta = gemm(A, B)
        run(program)
        C = meta["outputs"][0].numpy()
        err = np.linalg.norm(C - g["C"]) / np.linalg.norm(g["C"])
        print(os.path.basename(f), "rel err vs reference", err)
        assert err < 1e-10


def big(n, b, streams=4, profile=False, reps=2):
    dev = torch.device("cuda:0")
    nb = n // b
    X = [torch.empty(b, 128, dtype=torch.float64, device=dev) for _ in range(nb)]
    for j in range(nb):
        kernels.fill_random(X[j], seed=1234, row0=j * b, col0=0)
    for rep in range(reps):
        A = BigMatrix(f"big_{n}_{b}", shape=(n, n), shard_sizes=(b, b))
        A.free()
        for j in range(nb):
            for k in range(j + 1):
                t = torch.empty(b, b, dtype=torch.float64, device=dev)
                kernels._gemm_into(t, None, X[j], X[k], False, True, 1.0, 0.0)
                if j == k:
                    kernels.add_diag(t, float(n))
                A._put_block_ref(t, j, k)
        torch.cuda.synchronize()
        program, meta = cholesky(A)
        t0 = time.time()
        nn = len(program.program.nodes)
        t_expand = time.time() - t0
        program.start()
        torch.cuda.synchronize()
        t0 = time.time()
        job_runner.lambdapack_run(program, timeout=600, streams=streams, profile=profile, consume_inputs=True)
        torch.cuda.synchronize()
        dt = time.time() - t0
        print(f"N={n} b={b} streams={streams}: {nn} nodes, expand {t_expand:.2f}s, run {dt:.3f}s -> {n**3 / 3 / dt * 1e-12:.2f} TFLOP/s "
              f"status {program.program_status()} mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
        O = meta["outputs"][0]
        if rep == reps - 1:
            # residual check on a few tiles: (L L^T)_{jk} vs A_{jk}
            worst = 0.0
            for (j, k) in [(0, 0), (nb - 1, 0), (nb - 1, nb - 1), (nb // 2, nb // 3)]:
                acc = torch.zeros(b, b, dtype=torch.float64, device=dev)
                for i in range(k + 1):
                    acc += O._get_block_ref(j, i) @ O._get_block_ref(k, i).T
                ref = X[j] @ X[k].T
                if j == k:
                    ref += n * torch.eye(b, dtype=torch.float64, device=dev)
                worst = max(worst, float((acc - ref).norm() / ref.norm()))
            print("   residual ||LL^T - A||/||A|| on sampled tiles:", worst)
        if profile and rep == reps - 1:
            tl = job_runner.node_timeline(program)
            by = {}
            for name, vv, s, e, sid in tl:
                by.setdefault(name, []).append(e - s)
            for k2, v in by.items():
                print(f"   {k2}: n={len(v)} mean {np.mean(v):.3f} ms max {np.max(v):.3f} ms total {np.sum(v):.1f} ms")
            print("   makespan", max(e for _, _, _, e, _ in tl), "ms")
        meta["outputs"][0].free(); meta["intermediates"][0].free(); A.free()
        del program, meta, O


if __name__ == "__main__":
    golden_cholesky()
    golden_gemm()
    big(16384, 4096, reps=2)
    big(32768, 4096, reps=2, profile=True)
    big(65536, 4096, reps=2)
    big(65536, 4096, streams=8, reps=1)
    big(65536, 4096, streams=2, reps=1)
