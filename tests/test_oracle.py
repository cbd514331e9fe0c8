"""The CPU oracle (oracle/npw_oracle.py) against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py).  CPU only."""
import glob
import os

import numpy as np
import pytest

from oracle import npw_oracle as orc


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_kernels_match_reference(golden_dir):
    g = _load(golden_dir, "kernels.npz")
    assert np.array_equal(orc.syrk(g["s"], g["x"], g["y"]), g["syrk"])
    assert np.array_equal(orc.chol(g["spd"]), g["chol"])
    assert np.array_equal(orc.trsm(g["chol"], g["trsm_b"]), g["trsm"])
    assert np.array_equal(orc.add_matrices(g["p0"], g["p1"], g["p2"], g["p3"]), g["add"])
    assert np.array_equal(orc.gemm(g["ga"], g["gb"]), g["gemm"])
    assert np.array_equal(orc.mul(g["p0"], g["p1"]), g["mul"])


def test_trsm_is_right_solve_with_transposed_lower(golden_dir):
    g = _load(golden_dir, "kernels.npz")
    L = g["chol"]
    np.testing.assert_allclose(orc.trsm(L, g["trsm_b"]), g["trsm_b"] @ np.linalg.inv(L).T, rtol=1e-12, atol=1e-12)


def test_syrk_and_trsm_zero_shortcuts():
    s = np.ones((4, 4))
    assert orc.syrk(s, np.zeros((4, 3)), np.ones((4, 3))) is s           # kernels.py:213-214
    z = orc.trsm(np.eye(4), np.zeros((5, 4)))
    assert z.shape == (4, 5) and not z.any()                              # kernels.py:255-256 (shape quirk)


@pytest.mark.parametrize("name", ["cholesky_64_8", "cholesky_64_16", "cholesky_60_16", "cholesky_64_32_lam"])
def test_cholesky_program_matches_reference(golden_dir, name):
    g = _load(golden_dir, name + ".npz")
    n, b, lam = int(g["n"]), int(g["b"]), float(g["lambdav"])
    I = orc.OracleBigMatrix("A", (n, n), (b, b), lambdav=lam)
    orc.shard_matrix(I, g["A"])
    O, S = orc.run_cholesky(I)
    assert np.array_equal(O.numpy(), g["L"])
    s_keys = [k for k in g.files if k.startswith("S_")]
    assert len(s_keys) == len(S.store)
    for k in s_keys:
        i, j, kk = (int(x) for x in k.split("_")[1:])
        assert np.array_equal(S.store[(i, j, kk)].reshape(g[k].shape), g[k])
    # and the reference's own acceptance criterion (tests/test_alg_correctness.py:47-49)
    assert np.allclose(O.numpy(), np.linalg.cholesky(g["A"] + lam * np.eye(n)))


@pytest.mark.parametrize("name", ["gemm_64_16", "gemm_32_16"])
def test_gemm_program_matches_reference(golden_dir, name):
    g = _load(golden_dir, name + ".npz")
    n, b = int(g["n"]), int(g["b"])
    A = orc.OracleBigMatrix("A", (n, n), (b, b)); orc.shard_matrix(A, g["A"])
    B = orc.OracleBigMatrix("B", (n, n), (b, b)); orc.shard_matrix(B, g["B"])
    Out, _ = orc.run_gemm(A, B)
    assert np.array_equal(Out.numpy(), g["C"])
    assert np.allclose(orc.binops_gemm(A, B).numpy(), g["A"] @ g["B"])


@pytest.mark.parametrize("name", ["tsqr_256_32", "tsqr_128_16"])
def test_tsqr_program_matches_reference(golden_dir, name):
    g = _load(golden_dir, name + ".npz")
    m, b, nlev = int(g["m"]), int(g["b"]), int(g["nlev"])
    X = orc.OracleBigMatrix("X", (m, b), (b, b)); orc.shard_matrix(X, g["X"])
    Rs, Vs, Ts = orc.run_tsqr(X)
    R = Rs.get_block(nlev, 0)
    assert np.array_equal(R, g["R"])
    for k in g.files:
        if k[:2] in ("R_", "V_") or k.startswith("Tq_"):
            lvl, j = (int(x) for x in k.split("_")[1:])
            mat = {"R": Rs, "V": Vs, "Tq": Ts}[k.split("_")[0]]
            assert np.array_equal(mat.store[(lvl, j)], g[k]), k
    # reference acceptance criterion: R equals numpy's R up to row signs (tests/test_alg_correctness.py:95-102)
    Rnp = np.linalg.qr(g["X"])[1]
    np.testing.assert_allclose(np.abs(R), np.abs(Rnp), rtol=1e-10, atol=1e-12)


def test_qr_factor_is_compact_wy():
    rs = np.random.RandomState(3)
    a = rs.randn(96, 24)
    v, t, r = orc.qr_factor(a[:48], a[48:])
    q = np.eye(96) - v @ t @ v.T
    np.testing.assert_allclose(q.T @ q, np.eye(96), atol=1e-13)
    np.testing.assert_allclose((q @ np.vstack([r, np.zeros((72, 24))])), a, atol=1e-12)
    assert np.allclose(np.tril(r, -1), 0) and np.allclose(np.tril(t, -1), 0)
    assert np.allclose(np.diag(v), 1) and np.allclose(np.triu(v, 1), 0)


def test_bigmatrix_semantics():
    m = orc.OracleBigMatrix("m", (200, 200), (101, 101), parent_fn=orc.constant_zeros, lambdav=3.0)
    assert m.num_blocks(0) == 2 and m.block_idx_to_real_idx((1, 1)) == ((101, 200), (101, 200))
    assert m.get_block(1, 0).shape == (99, 101) and not m.get_block(1, 0).any()
    d = m.get_block(1, 1)
    assert d.shape == (99, 99) and np.array_equal(np.diag(d), np.full(99, 3.0))   # lambdav on diagonal reads
    with pytest.raises(Exception):
        m.put_block(np.zeros((5, 5)), 0, 0)                                        # safe shape check
    x = np.arange(200 * 200, dtype=np.float64).reshape(200, 200)
    m2 = orc.OracleBigMatrix("m2", (200, 200), (101, 101))
    orc.shard_matrix(m2, x)
    assert np.array_equal(m2.numpy(), x)
    with pytest.raises(Exception):
        orc.OracleBigMatrix("m3", (4, 4), (2, 2)).get_block(0, 0)                  # no parent_fn, missing key


def test_spd_generator_is_well_conditioned():
    b, n = 32, 64
    A = np.block([[orc.spd_tile(j, k, b, n, width=16) for k in range(2)] for j in range(2)])
    assert np.allclose(A, A.T)
    w = np.linalg.eigvalsh(A)
    assert w.min() > 0 and w.max() / w.min() < 10


# --------------------------------------------------------------------------- QR / BDFAC
def test_qr_update_kernels_match_reference(golden_dir):
    g = _load(golden_dir, "qr_kernels.npz")
    for tag in ("s", "l"):
        v, t, r = orc.qr_factor_triangular(g[f"{tag}_r0"], g[f"{tag}_r1"])
        assert np.array_equal(v, g[f"{tag}_tri_v"]) and np.array_equal(v, np.eye(v.shape[0]))   # kernels.py:120-122
        assert np.array_equal(t, g[f"{tag}_tri_t"]) and np.array_equal(r, g[f"{tag}_tri_r"])
        assert np.array_equal(orc.qr_leaf(g[f"{tag}_vq"], g[f"{tag}_tq"], g[f"{tag}_a"]), g[f"{tag}_leaf"])
        s01, s11 = orc.qr_trailing_update(g[f"{tag}_vm"], g[f"{tag}_tm"], g[f"{tag}_s0"], g[f"{tag}_s1"])
        assert np.array_equal(s01, g[f"{tag}_s01"]) and np.array_equal(s11, g[f"{tag}_s11"])
        n = g[f"{tag}_a"].shape[0]
        vl, tl, ll = orc.lq_factor(g[f"{tag}_wide"][:, :n], g[f"{tag}_wide"][:, n:])
        assert np.array_equal(vl, g[f"{tag}_vl"]) and np.array_equal(tl, g[f"{tag}_tl"]) and np.array_equal(ll, g[f"{tag}_ll"])
        l01, l11 = orc.lq_trailing_update(vl, tl, g[f"{tag}_c0"], g[f"{tag}_c1"])
        assert np.array_equal(l01, g[f"{tag}_l01"]) and np.array_equal(l11, g[f"{tag}_l11"])
        assert np.array_equal(orc.lq_leaf(g[f"{tag}_vl1"], g[f"{tag}_tl1"], g[f"{tag}_c0"]), g[f"{tag}_lqleaf"])


def test_blocked_t_is_the_diagonal_of_the_compact_wy_t(golden_dir):
    """dtpqrt (nb = 32 < n) stores the 32x32 diagonal blocks of the n x n compact-WY T side by side in rows 0..31 —
    the identity numpywren_b200/qr.py uses to produce the reference's T from its own full T."""
    g = _load(golden_dir, "qr_kernels.npz")
    r0, r1 = g["l_r0"], g["l_r1"]
    n = r0.shape[0]
    _, t_ref, r_ref = orc.qr_factor_triangular(r0, r1)
    v2, t_full, r_h = orc.qr_factor_triangular(r0, r1, semantics="householder")
    assert not t_ref[32:].any()
    for k0 in range(0, n, 32):
        w = min(32, n - k0)
        np.testing.assert_allclose(t_ref[:w, k0:k0 + w], t_full[k0:k0 + w, k0:k0 + w], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(r_h, r_ref, rtol=1e-12, atol=1e-14)
    # the stacked general QR has the same reflectors: [I; V2], same T, same R
    vs, ts, rs_ = orc.qr_factor(r0, r1)
    np.testing.assert_allclose(vs[:n], np.eye(n), atol=0)
    np.testing.assert_allclose(vs[n:], v2, rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(ts, t_full, rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(rs_, r_h, rtol=1e-11, atol=1e-13)


def _check_tiles(g, mats):
    n = 0
    for k in g.files:
        name = next((m for m in mats if k.startswith(m + "_")), None)
        if name is None:
            continue
        idx = tuple(int(x) for x in k[len(name) + 1:].split("_"))
        got = mats[name].store[idx]
        assert got.shape == g[k].shape and np.array_equal(got, g[k]), k
        n += 1
    assert n == sum(len(m.store) for m in mats.values())
    return n


@pytest.mark.parametrize("name", ["qr_28_7", "qr_16_8", "qr_24_8"])
def test_qr_program_matches_reference(golden_dir, name):
    g = _load(golden_dir, name + ".npz")
    n, b = int(g["n"]), int(g["b"])
    A = orc.OracleBigMatrix("A", (n, n), (b, b)); orc.shard_matrix(A, g["X"])
    Rs, Vs, Ts, S = orc.run_qr(A)
    assert _check_tiles(g, {"Vs": Vs, "Ts": Ts, "Rs": Rs, "Ss": S}) > 0


@pytest.mark.parametrize("name", ["bdfac_16_4", "bdfac_16_4_trunc2", "bdfac_15_5"])
def test_bdfac_program_matches_reference(golden_dir, name):
    g = _load(golden_dir, name + ".npz")
    n, b, trunc = int(g["n"]), int(g["b"]), int(g["truncate"])
    A = orc.OracleBigMatrix("A", (n, n), (b, b)); orc.shard_matrix(A, g["X"])
    mats = orc.run_bdfac(A, truncate=trunc)
    assert _check_tiles(g, mats) > 0


@pytest.mark.parametrize("n,b", [(28, 7), (16, 8), (24, 8), (96, 48)])
def test_qr_householder_semantics_meet_the_reference_test(n, b):
    """tests/test_alg_correctness.py:160-187: the last diagonal block of R equals np.linalg.qr's up to row signs —
    and so does every other block row."""
    X = np.random.RandomState(n).randn(n, n)
    A = orc.OracleBigMatrix("A", (n, n), (b, b)); orc.shard_matrix(A, X)
    Rs, _, _, _ = orc.run_qr(A, semantics="householder")
    nb = n // b
    R = np.zeros((n, n))
    for i in range(nb):
        R[i * b:(i + 1) * b, i * b:(i + 1) * b] = Rs.get_block(i, i, 0)
        for k in range(i + 1, nb):
            R[i * b:(i + 1) * b, k * b:(k + 1) * b] = Rs.get_block(i, k, 0)
    Rnp = np.linalg.qr(X)[1]
    np.testing.assert_allclose(np.abs(R), np.abs(Rnp), rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("n,b", [(16, 4), (15, 5), (48, 8)])
def test_bdfac_householder_semantics_meet_the_reference_test(n, b):
    """tests/test_alg_correctness.py:262-270: the block-bidiagonal factor has the singular values of the input."""
    X = np.random.RandomState(n + 1).randn(n, n)
    A = orc.OracleBigMatrix("A", (n, n), (b, b)); orc.shard_matrix(A, X)
    m = orc.run_bdfac(A, semantics="householder")
    fac = orc.bdfac_assemble(m["R_QR"], m["L_LQ"], n, b)
    np.testing.assert_allclose(np.linalg.svd(fac, compute_uv=False), np.linalg.svd(X, compute_uv=False), rtol=1e-10, atol=1e-12)


def test_slow_qr_restatement_agrees_with_geqrt_on_the_leading_block():
    """kernels.py:67-84: for a wide matrix the reflectors (and dlarft's T) depend on the leading m x m block only."""
    a = np.random.RandomState(8).randn(20, 50)
    v, t, r = orc.slow_qr(a)
    v1, t1, r1 = orc.fast_qr(np.ascontiguousarray(a[:, :20]))
    np.testing.assert_allclose(v, v1, atol=1e-13)
    np.testing.assert_allclose(t, t1, atol=1e-13)
    np.testing.assert_allclose(r[:, :20], r1, atol=1e-13)
    q = np.eye(20) - v @ t @ v.T
    np.testing.assert_allclose(q @ r, a, atol=1e-12)
    assert r.shape == (20, 50) and not np.tril(r, -1).any()
