"""Worker for tests/test_multi_gpu.py::test_qr_and_bdfac_sharded_over_gpus: one process per GPU.  Runs the QR and
BDFAC programs with their 1-D cyclic placements (alg_wrappers._loose) and checks the reference tests' criteria."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from numpywren_b200 import alg_wrappers, job_runner, parallel, qr  # noqa: E402
from numpywren_b200 import lambdapack as lp  # noqa: E402
from numpywren_b200.matrix import BigMatrix  # noqa: E402
from numpywren_b200.matrix_init import shard_matrix  # noqa: E402
from oracle import npw_oracle as orc  # noqa: E402


def gather_tile(m, idx, grid):
    """Tile idx of m on every rank (broadcast from its owner)."""
    owner = grid.owner(m, idx)
    shape = parallel._tile_shape(m, idx)
    buf = torch.empty(shape, dtype=torch.float64, device="cuda")
    if owner == grid.rank:
        buf.copy_(m.get_block(*idx).reshape(shape))
    dist.broadcast(buf, owner)
    return buf.cpu().numpy()


def main():
    grid = parallel.init_from_env("nccl")
    qr.set_qr_semantics("householder")
    n, b = 512, 64
    nb = n // b
    X = np.random.RandomState(21).randn(n, n)
    A = BigMatrix("mgq_A", shape=(n, n), shard_sizes=(b, b))
    shard_matrix(A, X)
    program, meta = alg_wrappers.qr(A)
    program.start()
    job_runner.lambdapack_run(program, timeout=300)
    assert program.program_status() == lp.PS.SUCCESS
    Rs = meta["outputs"][0]
    R = np.zeros((n, n))
    for i in range(nb):
        for k in range(i, nb):
            R[i * b:(i + 1) * b, k * b:(k + 1) * b] = gather_tile(Rs, (i, k, 0), grid)
    err = np.abs(np.abs(R) - np.abs(np.linalg.qr(X)[1])).max()
    assert err < 1e-9, err
    sent = torch.tensor([program._engine.comm.bytes_sent], dtype=torch.int64, device="cuda")
    dist.all_reduce(sent)
    assert int(sent.item()) > 0
    if grid.rank == 0:
        print(f"qr {n}/{b}: world {grid.world} max | |R| - |R_np| | {err:.2e} nvlink bytes {int(sent.item())}")

    B = BigMatrix("mgq_B", shape=(n, n), shard_sizes=(b, b))
    shard_matrix(B, X)
    program, meta = alg_wrappers.bdfac(B)
    program.start()
    job_runner.lambdapack_run(program, timeout=300)
    assert program.program_status() == lp.PS.SUCCESS
    L, Rq = meta["outputs"]
    fac = orc.bdfac_assemble(Rq, L, n, b, get=lambda m, *idx: gather_tile(m, idx, grid))
    serr = np.abs(np.linalg.svd(fac, compute_uv=False) - np.linalg.svd(X, compute_uv=False)).max()
    assert serr < 1e-9, serr
    # legacy binops.gemm across GPUs (algs.GEMM_ACC on the engine; A / B tiles travel, C tiles accumulate in place)
    from numpywren_b200 import binops
    rs = np.random.RandomState(5)
    ga, gb = rs.randn(768, 640), rs.randn(640, 512)
    GA = BigMatrix("mgq_GA", shape=ga.shape, shard_sizes=(128, 128)); shard_matrix(GA, ga)
    GB = BigMatrix("mgq_GB", shape=gb.shape, shard_sizes=(128, 128)); shard_matrix(GB, gb)
    XY = binops.gemm(None, GA, GB)
    C = XY.numpy()                                  # collective gather
    gerr = np.linalg.norm(C - ga @ gb) / np.linalg.norm(ga @ gb)
    assert gerr < 1e-13, gerr
    dist.barrier()
    if grid.rank == 0:
        print("MULTI_GPU_QR_OK", grid.world, f"svd err {serr:.2e} gemm err {gerr:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
