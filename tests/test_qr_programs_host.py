"""QR / BDFAC programs and the QR/LQ update kernels' HOST logic without a GPU (SURVEY §8 f#1, f#3).

numpywren_b200/qr.py only sequences C-ABI calls; here the same Python code runs against tests/_hostlib.HostLib
(a NumPy test double of the entry points it uses), and the results are compared
  * with the golden tiles the UNMODIFIED reference produced ("reference" semantics), and
  * with the criteria of the reference's own tests ("householder" semantics).
The `-m gpu` twin of this file (tests/test_qr_programs_gpu.py) runs the identical checks on the CUDA kernels.
"""
import os

import numpy as np
import pytest
import torch

from numpywren_b200 import alg_wrappers, kernels, qr
from numpywren_b200.matrix import BigMatrix
from numpywren_b200.matrix_init import shard_matrix
from oracle import npw_oracle as orc
import _hostlib  # tests/_hostlib.py (tests/ is on sys.path under pytest's rootdir conftest)

T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
RTOL = 1e-10


def close(got, want, tol=RTOL):
    got = got.numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    assert got.shape == want.shape, (got.shape, want.shape)
    scale = max(1.0, np.abs(want).max())
    assert np.abs(got - want).max() <= tol * scale, np.abs(got - want).max()


@pytest.fixture
def host(monkeypatch):
    prev = qr.get_qr_semantics()
    lib = _hostlib.install(monkeypatch)
    yield lib
    qr.set_qr_semantics(prev)


@pytest.mark.parametrize("tag", ["s", "l"])
def test_update_kernels_reference_semantics(host, golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "qr_kernels.npz"))
    qr.set_qr_semantics("reference")
    v, t, r = kernels.qr_factor_triangular(T(g[f"{tag}_r0"]), T(g[f"{tag}_r1"]))
    close(v, g[f"{tag}_tri_v"]); close(t, g[f"{tag}_tri_t"]); close(r, g[f"{tag}_tri_r"])
    close(kernels.qr_leaf(T(g[f"{tag}_vq"]), T(g[f"{tag}_tq"]), T(g[f"{tag}_a"])), g[f"{tag}_leaf"])
    s01, s11 = kernels.qr_trailing_update(T(g[f"{tag}_vm"]), T(g[f"{tag}_tm"]), T(g[f"{tag}_s0"]), T(g[f"{tag}_s1"]))
    close(s01, g[f"{tag}_s01"]); close(s11, g[f"{tag}_s11"])
    n = g[f"{tag}_a"].shape[0]
    wide = g[f"{tag}_wide"]
    vl, tl, ll = kernels.lq_factor(T(wide[:, :n]), T(wide[:, n:]))
    close(vl, g[f"{tag}_vl"]); close(tl, g[f"{tag}_tl"]); close(ll, g[f"{tag}_ll"])
    l01, l11 = kernels.lq_trailing_update(vl, tl, T(g[f"{tag}_c0"]), T(g[f"{tag}_c1"]))
    close(l01, g[f"{tag}_l01"]); close(l11, g[f"{tag}_l11"])
    close(kernels.lq_leaf(T(g[f"{tag}_vl1"]), T(g[f"{tag}_tl1"]), T(g[f"{tag}_c0"])), g[f"{tag}_lqleaf"])


def test_update_kernels_accept_strided_and_transposed_operands(host):
    """Views (column slices, .T) must reach the C-ABI with the right leading dimension / transpose flag."""
    rs = np.random.RandomState(5)
    qr.set_qr_semantics("householder")
    big = rs.randn(300, 400)
    a = T(big)[10:266, 7:263]                 # 256 x 256, ld 400
    v, t, r = kernels.qr_factor(a)
    vo, to, ro = orc.qr_factor(big[10:266, 7:263])
    close(v, vo); close(t, to); close(r, ro)
    s0 = T(rs.randn(384, 256)).T              # stored transposed
    want = orc.qr_leaf(vo, to, s0.numpy(), "householder")
    close(kernels.qr_leaf(v, t, s0), want)
    kinds = {c[0] for c in host.calls}
    assert "gemm" in kinds and "geqrt" in kinds
    # large products are issued in the DMMA core's NT form only (transA = 0, transB = 1)
    assert all((c[4], c[5]) == (0, 1) for c in host.calls if c[0] == "gemm" and min(c[1], c[2]) >= 128 and c[3] >= 64)


def test_triangular_merge_householder_semantics(host):
    rs = np.random.RandomState(6)
    qr.set_qr_semantics("householder")
    for n in (8, 40, 130):
        r0, r1 = np.triu(rs.randn(n, n)), np.triu(rs.randn(n, n))
        junk0 = r0 + np.tril(rs.randn(n, n), -1)      # dtpqrt never reads below the diagonals
        junk1 = r1 + np.tril(rs.randn(n, n), -1)
        v2, t, r = kernels.qr_factor_triangular(T(junk0), T(junk1))
        vo, to, ro = orc.qr_factor_triangular(r0, r1, "householder")
        close(v2, vo); close(t, to); close(r, ro)
        V = np.vstack([np.eye(n), v2.numpy()])
        Q = np.eye(2 * n) - V @ t.numpy() @ V.T
        close(Q.T @ np.vstack([r0, r1]), np.vstack([r.numpy(), np.zeros((n, n))]), 1e-12)


def _bigmatrix(name, X, b):
    A = BigMatrix(name, shape=X.shape, shard_sizes=(b, b), device="cpu")
    A.free()
    shard_matrix(A, X)
    return A


def _stored(m):
    return {idx: t.numpy() for idx, t in m._blocks_store.items()}


def _check_against_golden(g, mats):
    n = 0
    for k in g.files:
        name = next((m for m in mats if k.startswith(m + "_")), None)
        if name is None:
            continue
        idx = tuple(int(x) for x in k[len(name) + 1:].split("_"))
        got = _stored(mats[name])[idx]
        close(got.reshape(g[k].shape), g[k])
        n += 1
    assert n == sum(len(m._blocks_store) for m in mats.values())
    return n


@pytest.mark.parametrize("name", ["qr_28_7", "qr_16_8", "qr_24_8"])
def test_qr_program_reference_semantics_matches_golden(host, golden_dir, unique_key, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    qr.set_qr_semantics("reference")
    A = _bigmatrix(unique_key("qrA"), g["X"], int(g["b"]))
    program, meta = alg_wrappers.qr(A)
    for m in meta["outputs"] + meta["intermediates"]:
        m.free()
    assert _hostlib.run_in_program_order(program) == int(g["nnodes"])
    Rs, Vs, Ts = meta["outputs"]
    assert _check_against_golden(g, {"Vs": Vs, "Ts": Ts, "Rs": Rs, "Ss": meta["intermediates"][0]}) > 0


@pytest.mark.parametrize("name", ["bdfac_16_4", "bdfac_16_4_trunc2", "bdfac_15_5"])
def test_bdfac_program_reference_semantics_matches_golden(host, golden_dir, unique_key, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    qr.set_qr_semantics("reference")
    A = _bigmatrix(unique_key("bdA"), g["X"], int(g["b"]))
    program, meta = alg_wrappers.bdfac(A, truncate=int(g["truncate"]))
    mats = dict(zip(["L_LQ", "R_QR", "S_LQ", "S_QR", "T_QR", "V_QR", "V_LQ", "T_LQ"], meta["outputs"] + meta["intermediates"]))
    for m in mats.values():
        m.free()
    assert _hostlib.run_in_program_order(program) == int(g["nnodes"])
    assert _check_against_golden(g, mats) > 0


@pytest.mark.parametrize("n,b", [(28, 7), (16, 8), (192, 64)])
def test_qr_program_householder_semantics_meets_reference_test(host, unique_key, n, b):
    """tests/test_alg_correctness.py:160-187 (R equals np.linalg.qr's up to row signs), every block row."""
    qr.set_qr_semantics("householder")
    X = np.random.RandomState(n).randn(n, n)
    A = _bigmatrix(unique_key("qrA"), X, b)
    program, meta = alg_wrappers.qr(A)
    for m in meta["outputs"] + meta["intermediates"]:
        m.free()
    _hostlib.run_in_program_order(program)
    Rs = meta["outputs"][0]
    nb = n // b
    R = np.zeros((n, n))
    for i in range(nb):
        for k in range(i, nb):
            R[i * b:(i + 1) * b, k * b:(k + 1) * b] = Rs.get_block(i, k, 0).numpy()
    close(np.abs(R), np.abs(np.linalg.qr(X)[1]), 1e-9)
    # and tile for tile against the oracle's householder restatement
    Ao = orc.OracleBigMatrix("A", (n, n), (b, b)); orc.shard_matrix(Ao, X)
    Ro = orc.run_qr(Ao, semantics="householder")[0]
    for idx, t in _stored(Rs).items():
        close(t.reshape(Ro.store[idx].shape), Ro.store[idx], 1e-9)


@pytest.mark.parametrize("n,b", [(16, 4), (15, 5), (128, 32)])
def test_bdfac_program_householder_semantics_meets_reference_test(host, unique_key, n, b):
    """tests/test_alg_correctness.py:262-270: the block-bidiagonal factor keeps the singular values."""
    qr.set_qr_semantics("householder")
    X = np.random.RandomState(n + 1).randn(n, n)
    A = _bigmatrix(unique_key("bdA"), X, b)
    program, meta = alg_wrappers.bdfac(A)
    for m in meta["outputs"] + meta["intermediates"]:
        m.free()
    _hostlib.run_in_program_order(program)
    L, R = meta["outputs"]
    fac = orc.bdfac_assemble(R, L, n, b, get=lambda m, *idx: m.get_block(*idx).numpy())
    close(np.linalg.svd(fac, compute_uv=False), np.linalg.svd(X, compute_uv=False), 1e-10)


def test_gemm_kloop_program_matches_the_legacy_binops_schedule(host, unique_key):
    """algs.GEMM_ACC (the multi-GPU path of binops.gemm) through the host test double: same result as the oracle's
    restatement of binops._gemm_remote_0 (serial accumulation over the reduction index), ragged tiles included."""
    rs = np.random.RandomState(12)
    a, b = rs.randn(300, 260), rs.randn(260, 200)
    A = _bigmatrix(unique_key("gkA"), a, 128)
    B = _bigmatrix(unique_key("gkB"), b, 128)
    program, meta = alg_wrappers.gemm_kloop(A, B)
    for m in meta["outputs"] + meta["intermediates"]:
        m.free()
    cp = program.program
    assert len(cp.nodes) == 3 * 2 * (3 + 1)            # per output tile: 1 gemm + (K-1) gemm_acc + 1 identity
    for n in cp.nodes:                                   # every partial sum has one reader: in-place accumulation is legal
        if n.call.compute_name in ("gemm", "gemm_acc"):
            assert cp.num_readers(*n.writes[0]) == 1
    _hostlib.run_in_program_order(program)
    C = meta["outputs"][0].numpy()
    Ao = orc.OracleBigMatrix("A", a.shape, (128, 128)); orc.shard_matrix(Ao, a)
    Bo = orc.OracleBigMatrix("B", b.shape, (128, 128)); orc.shard_matrix(Bo, b)
    close(C, orc.binops_gemm(Ao, Bo).numpy(), 1e-12)
    close(C, a @ b, 1e-12)
    # large tiles go to the DMMA core in NT form after the transpose kernel re-lays B out; accumulation aliases C0 = C
    assert any(c[0] == "copy2d" and c[3] == 1 for c in host.calls)


@pytest.mark.parametrize("lower", [False, True])
@pytest.mark.parametrize("right", [True, False])
def test_trsm_flag_combinations_map_onto_the_one_native_solve(host, lower, right):
    """kernels.trsm(x, y, lower, right) = scipy.linalg.blas.dtrsm(1.0, x.T, y, lower=lower, side=int(right))
    (kernels.py:254-257) for all four flag combinations, with a FULL (non-triangular) x: BLAS reads one triangle only."""
    import scipy.linalg
    rs = np.random.RandomState(int(lower) * 2 + int(right))
    n, m = 24, 40
    x = rs.randn(n, n) + 6 * np.eye(n)
    y = rs.randn(m, n) if right else rs.randn(n, m)
    want = scipy.linalg.blas.dtrsm(1.0, x.T, y, lower=int(lower), side=int(right))
    got = kernels.trsm(T(x), T(y), lower=lower, right=right)
    close(got, np.ascontiguousarray(want), 1e-12)
    assert [c for c in host.calls if c[0] == "trsm_rlt"]            # all of them end in npw_trsm_rlt_f64


@pytest.mark.parametrize("m,n", [(12, 30), (130, 400)])
def test_qr_factor_wide_input_takes_the_slow_qr_path(host, m, n):
    """n > m: reference fast_qr falls back to slow_qr (dgeqrf + dlarft, kernels.py:67-84,94-95): v m x m, t m x m,
    r m x n upper trapezoidal."""
    a = np.random.RandomState(m).randn(m, n)
    v, t, r = kernels.qr_factor(T(a))
    vo, to, ro = orc.qr_factor(a)
    assert tuple(v.shape) == (m, m) and tuple(t.shape) == (m, m) and tuple(r.shape) == (m, n)
    close(v, vo, 1e-12); close(t, to, 1e-12); close(r, ro, 1e-12)
    q = np.eye(m) - vo @ to @ vo.T
    close(q @ ro, a, 1e-12)
