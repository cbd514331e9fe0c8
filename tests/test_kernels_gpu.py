"""Parity of every CUDA tile kernel (through the C-ABI) with the CPU oracle / reference golden vectors.
Tolerance: 1e-10 relative Frobenius error (BASELINE north_star), bit-exact for pure data movement / adds."""
import ctypes
import os

import numpy as np
import pytest
import torch

from numpywren_b200 import _capi, kernels
from oracle import npw_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-10


def dev(a, cuda_device):
    # np.array(...) copies: on the host harness (conftest NPW_B200_HOST_HARNESS) .to("cpu") would alias the ndarray
    return torch.from_numpy(np.array(a, dtype=np.float64, order="C")).to(cuda_device)


def rel(got, want):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else got
    return np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-300)


def test_extension_loaded_and_launch_counter(cuda_device):
    lib = _capi.load()
    before = _capi.launch_count()
    kernels.add_matrices(torch.ones(4, 4, dtype=torch.float64, device=cuda_device))
    torch.cuda.synchronize()
    assert _capi.launch_count() == before + 1
    assert lib.npw_last_error() is not None


def test_golden_kernel_vectors(golden_dir, cuda_device):
    g = np.load(os.path.join(golden_dir, "kernels.npz"))
    d = lambda k: dev(g[k], cuda_device)
    assert rel(kernels.syrk(d("s"), d("x"), d("y")), g["syrk"]) < TOL
    assert rel(kernels.chol(d("spd")), g["chol"]) < TOL
    assert rel(kernels.trsm(d("chol"), d("trsm_b")), g["trsm"]) < TOL
    assert rel(kernels.gemm(d("ga"), d("gb")), g["gemm"]) < TOL
    assert np.array_equal(kernels.add_matrices(d("p0"), d("p1"), d("p2"), d("p3")).cpu().numpy(), g["add"])  # bit exact
    assert np.array_equal(kernels.mul(d("p0"), d("p1")).cpu().numpy(), g["mul"])
    x = d("p0")
    assert kernels.identity(x) is x


@pytest.mark.parametrize("m,n,k", [(128, 128, 16), (128, 128, 128), (256, 384, 200), (100, 70, 36), (64, 64, 64), (33, 17, 5),
                                   (1, 1, 1), (130, 258, 1000), (512, 512, 2048), (96, 4096, 32), (4096, 96, 32)])
def test_syrk_shapes(cuda_device, m, n, k):
    rs = np.random.RandomState(m * 7 + n * 3 + k)
    s, x, y = rs.randn(m, n), rs.randn(m, k), rs.randn(n, k)
    got = kernels.syrk(dev(s, cuda_device), dev(x, cuda_device), dev(y, cuda_device))
    assert rel(got, orc.syrk(s, x, y)) < TOL


def test_syrk_tma_and_generic_paths_agree_on_unaligned_operands(cuda_device):
    """Odd leading dimensions / misaligned bases cannot be described by a TMA tensor map: the generic kernel runs."""
    rs = np.random.RandomState(0)
    big = dev(rs.randn(300, 301), cuda_device)       # ld = 301 (odd)
    x = big[:256, 1:129]                              # base misaligned by 8 bytes, ld odd
    y = big[20:276, 3:131]
    s = dev(rs.randn(256, 256), cuda_device)
    got = kernels.syrk(s, x, y)
    want = orc.syrk(s.cpu().numpy(), x.cpu().numpy(), y.cpu().numpy())
    assert rel(got, want) < TOL


def test_syrk_in_place_and_views(cuda_device):
    rs = np.random.RandomState(1)
    s, x, y = rs.randn(256, 256), rs.randn(256, 64), rs.randn(256, 64)
    st = dev(s, cuda_device)
    out = kernels.syrk(st, dev(x, cuda_device), dev(y, cuda_device), out=st)
    assert out is st and rel(st, orc.syrk(s, x, y)) < TOL
    # operands handed over as transposed (column-major) views, as BigMatrixView.T would
    xt = dev(x.T.copy(), cuda_device).T
    yt = dev(y.T.copy(), cuda_device).T
    assert rel(kernels.syrk(dev(s, cuda_device), xt, yt), orc.syrk(s, x, y)) < TOL


def test_syrk_zero_operand_matches_shortcut(cuda_device):
    # kernels.py:213-214 returns s when x or y is ~0; the kernel computes s - 0 = s exactly
    s = np.random.RandomState(2).randn(128, 128)
    got = kernels.syrk(dev(s, cuda_device), torch.zeros(128, 32, dtype=torch.float64, device=cuda_device),
                       dev(np.ones((128, 32)), cuda_device))
    assert np.array_equal(got.cpu().numpy(), s)


def test_syrk_empty(cuda_device):
    e = torch.zeros(0, 0, dtype=torch.float64, device=cuda_device)
    assert kernels.syrk(e, torch.zeros(0, 5, dtype=torch.float64, device=cuda_device),
                        torch.zeros(0, 5, dtype=torch.float64, device=cuda_device)).shape == (0, 0)
    s = dev(np.ones((8, 8)), cuda_device)
    z = torch.zeros(8, 0, dtype=torch.float64, device=cuda_device)
    assert np.array_equal(kernels.syrk(s, z, z).cpu().numpy(), np.ones((8, 8)))   # k = 0: s - 0


@pytest.mark.parametrize("m,n,k,ta,tb", [(64, 64, 64, 0, 0), (200, 136, 72, 0, 0), (200, 136, 72, 1, 0), (200, 136, 72, 0, 1),
                                         (200, 136, 72, 1, 1), (512, 384, 256, 0, 0), (1024, 1024, 512, 0, 0)])
def test_gemm_all_transpositions(cuda_device, m, n, k, ta, tb):
    rs = np.random.RandomState(m + n + k + ta + 2 * tb)
    a = rs.randn(k, m) if ta else rs.randn(m, k)
    b = rs.randn(n, k) if tb else rs.randn(k, n)
    got = kernels.gemm(dev(a, cuda_device), dev(b, cuda_device), transpose_A=bool(ta), transpose_B=bool(tb))
    assert rel(got, orc.gemm(a, b, bool(ta), bool(tb))) < TOL


def test_gemm_shape_mismatch_raises(cuda_device):
    with pytest.raises(ValueError):
        kernels.gemm(torch.zeros(4, 5, dtype=torch.float64, device=cuda_device),
                     torch.zeros(4, 5, dtype=torch.float64, device=cuda_device))


@pytest.mark.parametrize("n", [1, 7, 8, 64, 127, 128, 129, 200, 256, 1000, 2048])
def test_chol_and_trsm_sizes(cuda_device, n):
    rs = np.random.RandomState(n)
    x = rs.randn(n, n + 8)
    a = x @ x.T + n * np.eye(n)
    L = kernels.chol(dev(a, cuda_device))
    Lref = orc.chol(a)
    assert rel(L, Lref) < TOL
    assert not np.triu(L.cpu().numpy(), 1).any()                  # strict upper zeroed like np.linalg.cholesky
    m = max(n // 2, 3)
    b = rs.randn(m, n)
    got = kernels.trsm(dev(Lref, cuda_device), dev(b, cuda_device))
    assert rel(got, orc.trsm(Lref, b)) < TOL
    # the scheduler's path: inverted diagonal blocks from chol_async reused by trsm, result written in place
    Lq, info, inv = kernels.chol_async(dev(a, cuda_device))
    bt = dev(b, cuda_device)
    out = kernels.trsm_with_inverse(Lq, bt, inv, out=bt)
    assert out is bt and int(info.item()) == 0 and rel(bt, orc.trsm(Lref, b)) < TOL


def test_chol_reads_only_lower_triangle(cuda_device):
    rs = np.random.RandomState(5)
    x = rs.randn(96, 100)
    a = x @ x.T + 96 * np.eye(96)
    junk = np.tril(a) + np.triu(rs.randn(96, 96), 1) * 1e6       # syrk leaves garbage above the diagonal of S[i,i,i]
    assert rel(kernels.chol(dev(junk, cuda_device)), orc.chol(a)) < TOL


def test_chol_not_positive_definite_raises_linalgerror(cuda_device):
    a = np.eye(200)
    a[150, 150] = -1.0
    with pytest.raises(np.linalg.LinAlgError):
        kernels.chol(dev(a, cuda_device))
    L, info, _ = kernels.chol_async(dev(a, cuda_device))
    assert int(info.item()) == 151                                 # LAPACK INFO: first bad leading minor (1-based)
    with pytest.raises(np.linalg.LinAlgError):
        kernels.chol(torch.zeros(4, 5, dtype=torch.float64, device=cuda_device))


@pytest.mark.parametrize("lower", [False, True])
@pytest.mark.parametrize("right", [True, False])
def test_trsm_all_flag_combinations_match_blas(cuda_device, lower, right):
    """kernels.trsm(x, y, lower, right) = dtrsm(1.0, x.T, y, lower, side=right) (kernels.py:254-257); the three forms the
    DSL never uses are mapped onto the one native solve by transposes and index reversal."""
    import scipy.linalg
    rs = np.random.RandomState(int(lower) * 2 + int(right))
    n, m = 200, 136
    x = rs.randn(n, n) / n + np.eye(n)                      # well conditioned; FULL matrix: only one triangle may be read
    y = rs.randn(m, n) if right else rs.randn(n, m)
    want = np.ascontiguousarray(scipy.linalg.blas.dtrsm(1.0, x.T, y, lower=int(lower), side=int(right)))
    got = kernels.trsm(dev(x, cuda_device), dev(y, cuda_device), lower=lower, right=right)
    assert rel(got, want) < TOL


def test_trsm_ill_conditioned_factor_still_within_tolerance(cuda_device):
    rs = np.random.RandomState(9)
    n = 384
    L = np.tril(rs.randn(n, n)) + np.diag(np.linspace(1.0, 50.0, n))
    b = rs.randn(64, n)
    want = orc.trsm(L, b)
    got = kernels.trsm(dev(L, cuda_device), dev(b, cuda_device))
    resid = np.linalg.norm(got.cpu().numpy() @ L.T - b) / (np.linalg.norm(L) * np.linalg.norm(want))
    assert resid < 1e-13 and rel(got, want) < 1e-8


@pytest.mark.parametrize("count", [1, 2, 3, 4, 5, 8, 11])
def test_add_matrices_bit_exact(cuda_device, count):
    rs = np.random.RandomState(count)
    parts = [rs.randn(77, 33) for _ in range(count)]
    got = kernels.add_matrices(*[dev(p, cuda_device) for p in parts])
    assert np.array_equal(got.cpu().numpy(), orc.add_matrices(*parts))


def test_copy_transpose_diag_fill(cuda_device):
    lib = _capi.load()
    rs = np.random.RandomState(3)
    a = rs.randn(70, 45)
    t = dev(a, cuda_device)
    assert np.array_equal(kernels.transpose(t).cpu().numpy(), a.T)
    assert np.array_equal(kernels.transpose(t.T).cpu().numpy(), a)
    sq = dev(rs.randn(9, 9), cuda_device)
    ref = sq.cpu().numpy().copy()
    kernels.add_diag(sq, 2.5)
    ref[np.diag_indices(9)] += 2.5
    assert np.array_equal(sq.cpu().numpy(), ref)
    for mode, fn in ((1, np.triu), (2, np.tril)):
        w = dev(a[:45, :45], cuda_device)
        _capi.check(lib.npw_fill2d_f64(w.data_ptr(), 45, 45, 45, mode, 0.0, None), "fill")
        torch.cuda.synchronize()
        assert np.array_equal(w.cpu().numpy(), fn(a[:45, :45]))


def test_fill_random_is_reproducible_and_tile_consistent(cuda_device):
    full = torch.empty(64, 48, dtype=torch.float64, device=cuda_device)
    kernels.fill_random(full, seed=42)
    part = torch.empty(16, 8, dtype=torch.float64, device=cuda_device)
    kernels.fill_random(part, seed=42, row0=32, col0=40)
    assert torch.equal(part, full[32:48, 40:48])
    assert float(full.abs().max()) < 1.0 and abs(float(full.mean())) < 0.1


def test_benchmark_tile_4096(cuda_device):
    """The benchmark's tile size: parity of the three Cholesky kernels at b = 4096 through size-independent properties
    (the CPU oracle would need ~seconds per kernel here, still affordable: compare directly for syrk)."""
    b = 4096
    g = torch.Generator(device="cpu").manual_seed(0)
    x = torch.randn(b, 256, dtype=torch.float64, generator=g)
    y = torch.randn(b, 256, dtype=torch.float64, generator=g)
    s = torch.randn(b, b, dtype=torch.float64, generator=g)
    got = kernels.syrk(s.to(cuda_device), x.to(cuda_device), y.to(cuda_device))
    assert rel(got, orc.syrk(s.numpy(), x.numpy(), y.numpy())) < TOL
    xx = torch.randn(b, 128, dtype=torch.float64, generator=g).to(cuda_device)
    a = torch.empty(b, b, dtype=torch.float64, device=cuda_device)
    kernels._gemm_into(a, None, xx, xx, False, True, 1.0, 0.0)
    kernels.add_diag(a, float(b))
    L, info, inv = kernels.chol_async(a)
    assert int(info.item()) == 0
    recon = torch.empty_like(a)
    kernels._gemm_into(recon, None, L, L, False, True, 1.0, 0.0)
    assert float((recon - a).norm() / a.norm()) < 1e-14           # L L^T = A
    bmat = s.to(cuda_device)
    X = kernels.trsm_with_inverse(L, bmat, inv)
    back = torch.empty_like(a)
    kernels._gemm_into(back, None, X, L, False, True, 1.0, 0.0)                  # X L^T = B
    assert float((back - bmat).norm() / bmat.norm()) < 1e-13


@pytest.mark.parametrize("m,n", [(32, 32), (64, 32), (48, 17), (256, 32), (300, 70), (1024, 128), (4096, 64), (2048, 512)])
def test_qr_factor_matches_oracle(cuda_device, m, n):
    """V, T, R of the compact-WY QR against the oracle's LAPACK dgeqrt restatement of kernels.fast_qr."""
    rs = np.random.RandomState(m + n)
    a = rs.randn(m, n)
    V, T, R = kernels.qr_factor(dev(a, cuda_device))
    v, t, r = orc.qr_factor(a)
    assert rel(R, r) < TOL and rel(V, v) < TOL and rel(T, t) < TOL
    Vn, Tn, Rn = V.cpu().numpy(), T.cpu().numpy(), R.cpu().numpy()
    assert not np.tril(Rn, -1).any() and not np.tril(Tn, -1).any()
    assert np.array_equal(np.diag(Vn), np.ones(n)) and not np.triu(Vn, 1).any()
    q = np.eye(m) - Vn @ Tn @ Vn.T
    assert np.linalg.norm(q.T @ q - np.eye(m)) < 1e-11 * m
    assert np.linalg.norm(q[:, :n] @ Rn - a) / np.linalg.norm(a) < 1e-13


@pytest.mark.parametrize("m,n", [(70000, 64), (120000, 32), (66000, 40)])
def test_qr_factor_tall_tiles_use_the_other_panel_variants(cuda_device, m, n):
    """More than 448 rows per CTA (m > 66304 on 148 SMs) takes the shared-memory panel kernel, more than 672 the global
    one; 66000 rows is the register-resident kernel at the edge of its capacity.  Same V, T, R as LAPACK in all three."""
    a = np.random.RandomState(m % 1000 + n).randn(m, n)
    V, T, R = kernels.qr_factor(dev(a, cuda_device))
    v, t, r = orc.qr_factor(a)
    assert rel(R, r) < TOL and rel(V, v) < TOL and rel(T, t) < TOL
    assert not bool(torch.isnan(T).any())


def test_qr_factor_stacks_blocks_like_tsqr_merge(cuda_device):
    """TSQR merge node: qr_factor(R_left, R_right) on two upper-triangular factors (algs.py:36)."""
    rs = np.random.RandomState(12)
    r0, r1 = np.triu(rs.randn(32, 32)), np.triu(rs.randn(32, 32))
    V, T, R = kernels.qr_factor(dev(r0, cuda_device), dev(r1, cuda_device))
    v, t, r = orc.qr_factor(r0, r1)
    assert V.shape == (64, 32) and rel(R, r) < TOL and rel(V, v) < TOL and rel(T, t) < TOL


def test_qr_factor_rank_deficient_column(cuda_device):
    """A column that is already zero below the diagonal gives tau = 0 (H = I), as LAPACK's dlarfg does."""
    a = np.random.RandomState(13).randn(64, 8)
    a[1:, 0] = 0.0
    V, T, R = kernels.qr_factor(dev(a, cuda_device))
    v, t, r = orc.qr_factor(a)
    assert rel(R, r) < TOL and rel(V, v) < TOL and rel(T, t) < TOL and float(T[0, 0]) == 0.0
