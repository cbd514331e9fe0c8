"""Worker for tests/test_multi_gpu.py: one process per GPU (NCCL).  Factorises a golden fixture and a mid-size SPD
matrix with tiles block-cyclically sharded over the ranks and compares with the reference / oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from numpywren_b200 import job_runner, parallel  # noqa: E402
from numpywren_b200 import lambdapack as lp  # noqa: E402
from numpywren_b200.alg_wrappers import cholesky  # noqa: E402
from numpywren_b200.matrix import BigMatrix  # noqa: E402
from numpywren_b200.matrix_init import shard_matrix  # noqa: E402


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def main():
    grid = parallel.init_from_env("nccl")
    for name in ("cholesky_64_8", "cholesky_60_16", "cholesky_64_32_lam"):
        g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        n, b, lam = int(g["n"]), int(g["b"]), float(g["lambdav"])
        A = BigMatrix("mg_" + name, shape=(n, n), shard_sizes=(b, b), lambdav=lam)
        shard_matrix(A, g["A"])
        assert all(grid.is_mine(A, bi) for bi in A.block_idxs_exist)
        program, meta = cholesky(A)
        program.start()
        job_runner.lambdapack_run(program, timeout=120)
        assert program.program_status() == lp.PS.SUCCESS
        L = meta["outputs"][0].numpy()             # collective gather
        err = rel(L, g["L"])
        assert err < 1e-10, (name, err)
        eng = program._engine
        sent = torch.tensor([eng.comm.bytes_sent], dtype=torch.int64, device="cuda")
        dist.all_reduce(sent)
        if grid.rank == 0:
            print(f"{name}: world {grid.world} rel err {err:.2e} nvlink bytes {int(sent.item())}")
        assert int(sent.item()) > 0
    # mid-size: 2048 with 256-tiles vs numpy cholesky of the same matrix
    n, b = 2048, 256
    rs = np.random.RandomState(11)
    x = rs.randn(n, 96)
    a = x @ x.T + n * np.eye(n)
    A = BigMatrix("mg_mid", shape=(n, n), shard_sizes=(b, b))
    shard_matrix(A, a)
    program, meta = cholesky(A)
    program.start()
    job_runner.lambdapack_run(program, timeout=120, consume_inputs=True)
    L = meta["outputs"][0].numpy()
    err = rel(L, np.linalg.cholesky(a))
    assert err < 1e-10, err
    # TSQR: leaves live on rank j mod world, only R factors of the tree cross GPUs
    from numpywren_b200.alg_wrappers import tsqr
    g = np.load(os.path.join(ROOT, "tests", "golden", "tsqr_256_32.npz"))
    X = BigMatrix("mg_tsqr", shape=(256, 32), shard_sizes=(32, 32))
    program, meta = tsqr(X)              # placement is attached by the wrapper before tiles are stored
    shard_matrix(X, g["X"])
    assert sorted(X.block_idxs_exist) == [(j, 0) for j in range(8) if j % grid.world == grid.rank]
    program.start()
    job_runner.lambdapack_run(program, timeout=120)
    assert program.program_status() == lp.PS.SUCCESS
    Rs = meta["outputs"][0]
    nlev = int(g["nlev"])
    if grid.owner(Rs, (nlev, 0)) == grid.rank:
        R = Rs.get_block(nlev, 0).cpu().numpy()
        terr = rel(R, g["R"])
        assert terr < 1e-10, terr
        print(f"tsqr_256_32: world {grid.world} rel err {terr:.2e}")
    # a non-SPD matrix must fail on EVERY rank, whichever rank owns the offending tile
    bad = np.eye(64)
    bad[50, 50] = -1.0
    B = BigMatrix("mg_bad", shape=(64, 64), shard_sizes=(16, 16))
    shard_matrix(B, bad)
    program, meta = cholesky(B)
    program.start()
    try:
        job_runner.lambdapack_run(program, timeout=60)
        raise SystemExit("expected LinAlgError")
    except np.linalg.LinAlgError:
        pass
    assert program.program_status() == lp.PS.EXCEPTION
    dist.barrier()
    if grid.rank == 0:
        print("MULTI_GPU_OK", grid.world, f"mid err {err:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
