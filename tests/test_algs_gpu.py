"""End-to-end parity of the LambdaPACK programs on the GPU engine (alg_wrappers + lambdapack_run, the
call pattern of reference tests/test_alg_correctness.py:31-50,137-156) with the reference's golden outputs
and the CPU oracle."""
import glob
import os

import numpy as np
import pytest
import torch

from numpywren_b200 import job_runner, kernels
from numpywren_b200 import lambdapack as lp
from numpywren_b200.alg_wrappers import cholesky, gemm
from numpywren_b200.matrix import BigMatrix
from numpywren_b200.matrix_init import shard_matrix
from oracle import npw_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rel(got, want):
    return np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-300)


def run(program, **kw):
    program.start()
    out = job_runner.lambdapack_run(program, timeout=120, **kw)
    assert program.program_status() == lp.PS.SUCCESS
    return out


@pytest.mark.parametrize("name", ["cholesky_64_8", "cholesky_64_16", "cholesky_60_16", "cholesky_64_32_lam"])
@pytest.mark.parametrize("inplace", [True, False])
def test_cholesky_golden(golden_dir, unique_key, cuda_device, name, inplace):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    n, b, lam = int(g["n"]), int(g["b"]), float(g["lambdav"])
    A = BigMatrix(unique_key(name), shape=(n, n), shard_sizes=(b, b), lambdav=lam, write_header=True)
    shard_matrix(A, g["A"])
    program, meta = cholesky(A)
    res = run(program, inplace=inplace)
    assert len(res["executed_messages"]) == int(g["nnodes"])
    L = meta["outputs"][0].numpy()
    assert rel(L, g["L"]) < TOL
    assert np.allclose(L, np.linalg.cholesky(g["A"] + lam * np.eye(n)))        # the reference test's own criterion
    assert np.array_equal(A.numpy(), g["A"] + lam * np.eye(n))                 # input tiles were not consumed
    if not inplace:
        # with aliasing off every SSA version of S survives, as in the reference's S3 store
        S = meta["intermediates"][0]
        for k in g.files:
            if k.startswith("S_"):
                i, j, kk = (int(x) for x in k.split("_")[1:])
                got = S.get_block(i, j, kk).cpu().numpy()
                ref = g[k].reshape(got.shape)
                if j == kk:   # diagonal tiles: only the lower triangle is ever consumed (chol reads 'L')
                    got, ref = np.tril(got), np.tril(ref)
                assert rel(got, ref) < TOL, k
    else:
        assert len(meta["intermediates"][0].block_idxs_exist) == 0             # every S buffer was re-used in place
    for m in meta["outputs"] + meta["intermediates"] + [A]:
        m.free()


def test_program_wait_runs_the_engine(golden_dir, unique_key, cuda_device):
    g = np.load(os.path.join(golden_dir, "cholesky_64_16.npz"))
    A = BigMatrix(unique_key("w"), shape=(64, 64), shard_sizes=(16, 16))
    shard_matrix(A, g["A"])
    program, meta = cholesky(A)
    program.start()
    program.wait()                       # no separate worker: the waiting thread drives the GPU
    assert program.program_status() == lp.PS.SUCCESS
    assert rel(meta["outputs"][0].numpy(), g["L"]) < TOL
    assert program.get_flops() > 0 and program.get_read() > 0 and program.get_write() > 0


def test_cholesky_truncate(unique_key, cuda_device):
    rs = np.random.RandomState(4)
    x = rs.randn(96, 96)
    a = x @ x.T + 96 * np.eye(96)
    A = BigMatrix(unique_key("tr"), shape=(96, 96), shard_sizes=(16, 16))
    shard_matrix(A, a)
    program, meta = cholesky(A, truncate=2)
    run(program)
    L = meta["outputs"][0].numpy()
    want = np.linalg.cholesky(a)
    assert rel(L[:64, :64], want[:64, :64]) < TOL and not L[64:].any()


def test_not_positive_definite_program_fails_loudly(unique_key, cuda_device):
    a = np.eye(64)
    a[40, 40] = -3.0
    A = BigMatrix(unique_key("bad"), shape=(64, 64), shard_sizes=(16, 16))
    shard_matrix(A, a)
    program, meta = cholesky(A)
    program.start()
    with pytest.raises(np.linalg.LinAlgError):
        job_runner.lambdapack_run(program, timeout=60)
    assert program.program_status() == lp.PS.EXCEPTION


@pytest.mark.parametrize("name", ["gemm_64_16", "gemm_32_16"])
def test_gemm_golden(golden_dir, unique_key, cuda_device, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    n, b = int(g["n"]), int(g["b"])
    A = BigMatrix(unique_key("ga"), shape=(n, n), shard_sizes=(b, b))
    B = BigMatrix(unique_key("gb"), shape=(n, n), shard_sizes=(b, b))
    shard_matrix(A, g["A"]); shard_matrix(B, g["B"])
    program, meta = gemm(A, B)
    run(program)
    assert rel(meta["outputs"][0].numpy(), g["C"]) < TOL


def test_cholesky_mid_size_against_oracle(unique_key, cuda_device):
    """N=2048 with 256-tiles (8x8 tile grid, 120 nodes): full oracle replay on the host vs the GPU engine, same tiles."""
    n, b = 2048, 256
    nb = n // b
    A = BigMatrix(unique_key("mid"), shape=(n, n), shard_sizes=(b, b))
    I = orc.OracleBigMatrix("I", (n, n), (b, b))
    for j in range(nb):
        for k in range(nb):
            t = orc.spd_tile(j, k, b, n, width=64)
            I.put_block(t, j, k)
            A.put_block(t, j, k)
    O_ref, _ = orc.run_cholesky(I)
    program, meta = cholesky(A)
    run(program, streams=3)
    assert rel(meta["outputs"][0].numpy(), O_ref.numpy()) < TOL


def test_cholesky_benchmark_tile_properties(unique_key, cuda_device):
    """Benchmark tile (4096) at N=16384: too big for the oracle in test time, so check the factorisation identity
    ||L L^T - A|| / ||A|| and lower-triangularity, plus agreement with cuSOLVER's potrf on the assembled matrix."""
    n, b = 16384, 4096
    nb = n // b
    X = [torch.empty(b, 128, dtype=torch.float64, device=cuda_device) for _ in range(nb)]
    for j in range(nb):
        kernels.fill_random(X[j], seed=7, row0=j * b)
    A = BigMatrix(unique_key("bt"), shape=(n, n), shard_sizes=(b, b))
    full = torch.empty(n, n, dtype=torch.float64, device=cuda_device)
    for j in range(nb):
        for k in range(nb):
            t = X[j] @ X[k].T
            if j == k:
                t += n * torch.eye(b, dtype=torch.float64, device=cuda_device)
            full[j * b:(j + 1) * b, k * b:(k + 1) * b] = t
            A._put_block_ref(t, j, k)
    program, meta = cholesky(A)
    run(program, consume_inputs=True)
    O = meta["outputs"][0]
    L = torch.zeros(n, n, dtype=torch.float64, device=cuda_device)
    for j in range(nb):
        for k in range(j + 1):
            L[j * b:(j + 1) * b, k * b:(k + 1) * b] = O._get_block_ref(j, k)
    assert float(torch.triu(L, 1).abs().max()) == 0.0
    resid = float((L @ L.T - full).norm() / full.norm())
    assert resid < 1e-14
    Lref = torch.linalg.cholesky(full)
    assert float((L - Lref).norm() / Lref.norm()) < TOL
    for m in (O, meta["intermediates"][0], A):
        m.free()


def test_profile_timeline(unique_key, cuda_device):
    rs = np.random.RandomState(8)
    x = rs.randn(64, 64)
    A = BigMatrix(unique_key("pf"), shape=(64, 64), shard_sizes=(16, 16))
    shard_matrix(A, x @ x.T + 64 * np.eye(64))
    program, _ = cholesky(A)
    program.start()
    job_runner.lambdapack_run(program, profile=True)
    tl = job_runner.node_timeline(program)
    assert len(tl) == 20 and {t[0] for t in tl} == {"chol", "trsm", "syrk"}
    assert all(e >= s for _, _, s, e, _ in tl)


@pytest.mark.parametrize("name", ["tsqr_256_32", "tsqr_128_16"])
def test_tsqr_golden(golden_dir, unique_key, cuda_device, name):
    """tests/test_alg_correctness.py:72-102: R of the TSQR tree vs the reference run (elementwise) and vs numpy up to row signs."""
    from numpywren_b200.alg_wrappers import tsqr
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    m, b, nlev = int(g["m"]), int(g["b"]), int(g["nlev"])
    X = BigMatrix(unique_key(name), shape=(m, b), shard_sizes=(b, b))
    shard_matrix(X, g["X"])
    program, meta = tsqr(X)
    res = run(program)
    assert len(res["executed_messages"]) == int(g["nnodes"])
    Rs, Vs, Ts = meta["outputs"]
    R = Rs.get_block(nlev, 0).cpu().numpy()
    assert rel(R, g["R"]) < TOL
    Rnp = np.linalg.qr(g["X"])[1]
    assert np.allclose(np.abs(R), np.abs(Rnp))
    for k in g.files:
        if k[:2] in ("R_", "V_") or k.startswith("Tq_"):
            lvl, j = (int(x) for x in k.split("_")[1:])
            mat = {"R": Rs, "V": Vs, "Tq": Ts}[k.split("_")[0]]
            assert rel(mat.get_block(lvl, j).cpu().numpy(), g[k]) < TOL, k


def test_async_upload_and_host_mirror(golden_dir, unique_key, cuda_device):
    """The end-to-end path bench.py times: pinned host tiles go in through put_block (asynchronous H2D, consumers wait on
    the tile's event), factor tiles are written through to pinned host memory while the program is still running."""
    g = np.load(os.path.join(golden_dir, "cholesky_64_16.npz"))
    n, b = 64, 16
    A = BigMatrix(unique_key("e2e"), shape=(n, n), shard_sizes=(b, b))
    for (j, k) in A.block_idxs:
        h = torch.from_numpy(np.ascontiguousarray(g["A"][j * b:(j + 1) * b, k * b:(k + 1) * b])).pin_memory()
        A.put_block(h, j, k, non_blocking=True)
        assert A._ready_event(j, k) is not None
    program, meta = cholesky(A)
    O = meta["outputs"][0]
    O.mirror_to_host()
    run(program, consume_inputs=True)
    host = O.wait_mirror()
    assert sorted(host) == [(j, k) for j in range(4) for k in range(j + 1)]
    L = np.zeros((n, n))
    for (j, k), t in host.items():
        assert t.is_pinned()
        L[j * b:(j + 1) * b, k * b:(k + 1) * b] = t.numpy()
    assert rel(L, g["L"]) < TOL
    assert rel(O.numpy(), g["L"]) < TOL


def test_binops_gemm(unique_key, cuda_device):
    """tests/test_gemm.py:12-41 (legacy binops.gemm): single shard X X^T through a transposed view, and 2x2 shards X Y;
    plus BASELINE config 1's shape class (tile 1024) against the oracle's restatement."""
    from numpywren_b200 import binops
    rs = np.random.RandomState(21)
    X = rs.randn(16, 16)
    Xs = BigMatrix(unique_key("bx"), shape=X.shape, shard_sizes=X.shape)
    shard_matrix(Xs, X)
    XXT = binops.gemm(None, Xs, Xs.T, Xs.bucket, 1)
    assert np.all(np.isclose(X.dot(X.T), XXT.numpy()))
    Y = rs.randn(16, 16)
    Xh = BigMatrix(unique_key("bx2"), shape=X.shape, shard_sizes=(8, 8)); shard_matrix(Xh, X)
    Yh = BigMatrix(unique_key("by2"), shape=Y.shape, shard_sizes=(8, 8)); shard_matrix(Yh, Y)
    XY = binops.gemm(None, Xh, Yh, Xh.bucket, 1)
    assert np.all(np.isclose(X.dot(Y), XY.numpy()))
    n, b = 2048, 512
    A, B = rs.randn(n, n), rs.randn(n, n)
    Am = BigMatrix(unique_key("ba"), shape=(n, n), shard_sizes=(b, b)); shard_matrix(Am, A)
    Bm = BigMatrix(unique_key("bb"), shape=(n, n), shard_sizes=(b, b)); shard_matrix(Bm, B)
    oa = orc.OracleBigMatrix("a", (n, n), (b, b)); orc.shard_matrix(oa, A)
    ob = orc.OracleBigMatrix("b", (n, n), (b, b)); orc.shard_matrix(ob, B)
    assert rel(binops.gemm(None, Am, Bm).numpy(), orc.binops_gemm(oa, ob).numpy()) < TOL
    with pytest.raises(Exception, match="shard size"):
        binops.gemm(None, Xs, Yh)
