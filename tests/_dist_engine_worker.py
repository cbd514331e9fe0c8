"""Worker for tests/test_distributed_cpu.py::test_engine_over_gloo: the REAL multi-rank engine on the host.

One process per rank (gloo).  tests/_fakecuda.py replaces CUDA streams/events by inert stand-ins and the C-ABI by the
NumPy double; NPW_B200_EXCHANGE=nccl selects the isend/irecv tile exchange, which gloo carries between CPU tensors.  Each
rank then walks the same DAG, runs the nodes it owns, ships the tiles other ranks need — exactly the code path of a
multi-GPU run minus the kernels and the NVLink copies — for Cholesky, QR, BDFAC and the distributed binops.gemm."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _fakecuda  # noqa: E402
from numpywren_b200 import alg_wrappers, binops, job_runner, parallel, qr  # noqa: E402
from numpywren_b200 import lambdapack as lp  # noqa: E402
from numpywren_b200.matrix import BigMatrix  # noqa: E402
from numpywren_b200.matrix_init import shard_matrix  # noqa: E402
from oracle import npw_oracle as orc  # noqa: E402


def gather_tile(m, idx, grid):
    owner = grid.owner(m, idx)
    shape = parallel._tile_shape(m, idx)
    buf = torch.empty(shape, dtype=torch.float64)
    if owner == grid.rank:
        buf.copy_(m.get_block(*idx).reshape(shape))
    dist.broadcast(buf, owner)
    return buf.numpy()


def main():
    os.environ["NPW_B200_EXCHANGE"] = "nccl"
    grid = parallel.init_from_env("gloo")
    mp = pytest.MonkeyPatch()
    _fakecuda.install(mp)
    mp.setattr(torch.cuda, "current_device", lambda: 0)
    try:
        # ---- Cholesky: golden fixture, block-cyclic ownership, panel tiles cross ranks
        g = np.load(os.path.join(ROOT, "tests", "golden", "cholesky_64_8.npz"))
        A = BigMatrix("de_chol", shape=(64, 64), shard_sizes=(8, 8), device="cpu")
        shard_matrix(A, g["A"])
        program, meta = alg_wrappers.cholesky(A)
        plan_s = job_runner.prepare(program, streams=8, consume_inputs=True)   # as bench.py does before its timed region
        assert plan_s >= 0 and program._engine.comm is not None and program._engine.n_streams == 8
        program.start()
        job_runner.lambdapack_run(program, timeout=120, streams=8, consume_inputs=True)
        assert program.program_status() == lp.PS.SUCCESS
        L = meta["outputs"][0].numpy()
        assert np.linalg.norm(L - g["L"]) / np.linalg.norm(g["L"]) < 1e-12
        sent = torch.tensor([program._engine.comm.bytes_sent], dtype=torch.int64)
        dist.all_reduce(sent)
        assert int(sent.item()) > 0

        # ---- QR (Householder semantics): R equals numpy's up to row signs, every block row
        qr.set_qr_semantics("householder")
        n, b = 48, 8
        nb = n // b
        X = np.random.RandomState(21).randn(n, n)
        Aq = BigMatrix("de_qr", shape=(n, n), shard_sizes=(b, b), device="cpu")
        shard_matrix(Aq, X)
        program, meta = alg_wrappers.qr(Aq)
        program.start()
        job_runner.lambdapack_run(program, timeout=120, free_intermediates=True)   # dead S versions are dropped per rank
        assert program.program_status() == lp.PS.SUCCESS
        freed = torch.tensor([program._engine.freed_tiles], dtype=torch.int64)
        dist.all_reduce(freed)
        assert int(freed.item()) > 0
        Rs = meta["outputs"][0]
        R = np.zeros((n, n))
        for i in range(nb):
            for k in range(i, nb):
                R[i * b:(i + 1) * b, k * b:(k + 1) * b] = gather_tile(Rs, (i, k, 0), grid)
        assert np.abs(np.abs(R) - np.abs(np.linalg.qr(X)[1])).max() < 1e-10

        # ---- BDFAC: the block-bidiagonal factor keeps the singular values
        Ab = BigMatrix("de_bd", shape=(n, n), shard_sizes=(b, b), device="cpu")
        shard_matrix(Ab, X)
        program, meta = alg_wrappers.bdfac(Ab)
        program.start()
        job_runner.lambdapack_run(program, timeout=120)
        assert program.program_status() == lp.PS.SUCCESS
        Lq, Rq = meta["outputs"]
        fac = orc.bdfac_assemble(Rq, Lq, n, b, get=lambda m, *idx: gather_tile(m, idx, grid))
        assert np.abs(np.linalg.svd(fac, compute_uv=False) - np.linalg.svd(X, compute_uv=False)).max() < 1e-10

        # ---- legacy binops.gemm across ranks (algs.GEMM_ACC on the engine)
        rs = np.random.RandomState(5)
        ga, gb = rs.randn(40, 56), rs.randn(56, 24)
        GA = BigMatrix("de_ga", shape=ga.shape, shard_sizes=(8, 8), device="cpu"); shard_matrix(GA, ga)
        GB = BigMatrix("de_gb", shape=gb.shape, shard_sizes=(8, 8), device="cpu"); shard_matrix(GB, gb)
        # the engine refuses CPU tiles in binops.gemm's single-GPU path; the distributed path goes through the engine
        C = binops.gemm(None, GA, GB).numpy()
        assert np.linalg.norm(C - ga @ gb) / np.linalg.norm(ga @ gb) < 1e-13

        # ---- TSQR: leaves live on rank j mod world, only the R factors of the tree cross ranks
        gt = np.load(os.path.join(ROOT, "tests", "golden", "tsqr_256_32.npz"))
        Xt = BigMatrix("de_tsqr", shape=(256, 32), shard_sizes=(32, 32), device="cpu")
        program, meta = alg_wrappers.tsqr(Xt)          # the wrapper attaches the placement before tiles are stored
        shard_matrix(Xt, gt["X"])
        assert sorted(Xt.block_idxs_exist) == [(j, 0) for j in range(8) if j % grid.world == grid.rank]
        program.start()
        job_runner.lambdapack_run(program, timeout=120)
        assert program.program_status() == lp.PS.SUCCESS
        Rt = meta["outputs"][0]
        nlev = int(gt["nlev"])
        Rfin = gather_tile(Rt, (nlev, 0), grid)
        assert np.linalg.norm(Rfin - gt["R"]) / np.linalg.norm(gt["R"]) < 1e-12

        # ---- a failure on one rank is a failure on all
        bad = np.eye(32)
        bad[20, 20] = -1.0
        Bm = BigMatrix("de_bad", shape=(32, 32), shard_sizes=(8, 8), device="cpu")
        shard_matrix(Bm, bad)
        program, meta = alg_wrappers.cholesky(Bm)
        program.start()
        try:
            job_runner.lambdapack_run(program, timeout=60)
            raise SystemExit("expected LinAlgError")
        except np.linalg.LinAlgError:
            pass
        dist.barrier()
        if grid.rank == 0:
            print("DIST_ENGINE_OK", grid.world)
    finally:
        mp.undo()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
