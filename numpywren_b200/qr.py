"""QR / LQ tile kernels on B200: compact-WY factorisations and the reflector-application updates of the
reference's TSQR, QR and BDFAC programs.

Reference: kernels.qr_factor (kernels.py:127-130) = fast_qr(np.vstack(blocks)) (kernels.py:86-105, LAPACK dgeqrt3),
qr_factor_triangular (:132-134 → fast_qr_triangular :107-124, LAPACK dtpqrt), lq_factor (:145-150), qr_leaf (:160-164),
lq_leaf (:154-157), qr_trailing_update (:181-188), lq_trailing_update (:199-208).

Every O(n^3) step runs through the DMMA GEMM core or the Householder panel kernels of libnpw_b200; this module only
sequences C-ABI calls on the current CUDA stream (no host synchronisation, no CPU arithmetic).

Two semantics (``set_qr_semantics`` / env NPW_B200_QR_SEMANTICS):

``"reference"`` (default) reproduces what the reference code computes, including its work-in-progress placeholders:
  * qr_leaf returns ``S0 - V.T @ S0`` (the compact-WY form is commented out at kernels.py:161-163);
  * fast_qr_triangular returns ``np.triu(x1.T).T`` of dtpqrt's (upper-triangular) V with a unit diagonal, i.e. the
    identity (kernels.py:120-122), and dtpqrt's T in LAPACK's blocked storage: with nb = min(n, 32) only the first nb
    rows are used, holding the 32x32 diagonal blocks of the compact-WY T side by side (kernels.py:117-119).
``"householder"`` is the mathematically intended algorithm, the one the reference's own tests assert
(tests/test_alg_correctness.py:160-187: R equals np.linalg.qr's up to row signs): qr_leaf applies
``(I - V T V^T)^T``, qr_factor_triangular returns the n x n upper-triangular lower half V2 of the reflectors
(Q = I - [I; V2] T [I; V2]^T) and the full n x n T.
"""
from __future__ import annotations

import os

import torch

from . import _capi
from .kernels import _check_tile, _gemm_any, _mat, _stream, add_diag, add_matrices, transpose

_SEMANTICS = os.environ.get("NPW_B200_QR_SEMANTICS", "reference")
_TPQRT_NB = 32  # kernels.py:118: nb = min(n, 32)


def set_qr_semantics(mode: str) -> str:
    """Select "reference" or "householder" (module docstring); returns the previous setting."""
    global _SEMANTICS
    if mode not in ("reference", "householder"):
        raise ValueError("qr semantics must be 'reference' or 'householder'")
    prev, _SEMANTICS = _SEMANTICS, mode
    return prev


def get_qr_semantics() -> str:
    return _SEMANTICS


def _copy_into(dst, src):
    """dst[:, :] = src through the library's strided copy / transpose kernel (dst row-major)."""
    lib = _capi.load()
    sm, ld, tr = _mat(src, "block")
    if tr:
        rc = lib.npw_copy2d_f64(dst.data_ptr(), max(1, dst.stride(0)), sm.data_ptr(), ld, src.shape[1], src.shape[0], 1, _stream())
    else:
        rc = lib.npw_copy2d_f64(dst.data_ptr(), max(1, dst.stride(0)), sm.data_ptr(), ld, src.shape[0], src.shape[1], 0, _stream())
    _capi.check(rc, "npw_copy2d_f64")


def _fill2d(t, mode, value=0.0):
    rc = _capi.load().npw_fill2d_f64(t.data_ptr(), max(1, t.stride(0)), t.shape[0], t.shape[1], int(mode), float(value), _stream())
    _capi.check(rc, "npw_fill2d_f64")
    return t


def _geqrt_inplace(V):
    """Compact-WY QR of the row-major m x n matrix in V (m >= n); V is overwritten by the reflectors → (V, T, R)."""
    lib = _capi.load()
    m, n = V.shape
    dev = V.device
    T = torch.empty((n, n), dtype=torch.float64, device=dev)
    R = torch.empty((n, n), dtype=torch.float64, device=dev)
    work = torch.empty(max(1, lib.npw_geqrt_work_bytes(m, n) // 8), dtype=torch.float64, device=dev)
    rc = lib.npw_geqrt_f64(V.data_ptr(), max(1, n), T.data_ptr(), max(1, n), R.data_ptr(), max(1, n), V.data_ptr(), max(1, n),
                           m, n, work.data_ptr(), _stream())
    _capi.check(rc, "npw_geqrt_f64")
    return V, T, R


def qr_factor(*blocks, **kwargs):
    if not blocks:
        raise TypeError("qr_factor expects at least one tile")
    for i, b in enumerate(blocks):
        _check_tile(b, f"blocks[{i}]")
    n = blocks[0].shape[1]
    for b in blocks[1:]:
        if b.shape[1] != n:
            raise ValueError("all the input array dimensions except for the concatenation axis must match exactly")
    m = sum(b.shape[0] for b in blocks)
    dev = blocks[0].device
    A = torch.empty((m, n), dtype=torch.float64, device=dev)
    # np.vstack: the stacked copy becomes the working matrix that V overwrites
    r0 = 0
    for b in blocks:
        _copy_into(A[r0:r0 + b.shape[0]], b)
        r0 += b.shape[0]
    if n <= m:
        return _geqrt_inplace(A)
    # wide input (reference fast_qr -> slow_qr, kernels.py:67-84,94-95: dgeqrf + dlarft): the m reflectors come from the
    # leading m x m block alone; the remaining columns only receive Q^T:  R = [R1 | A2 - V T^T (V^T A2)]
    V = torch.empty((m, m), dtype=torch.float64, device=dev)
    _copy_into(V, A[:, :m])
    V, T, R1 = _geqrt_inplace(V)
    A2 = A[:, m:]
    W = _gemm_any(_new(m, n - m, A), None, V, A2, True, False, 1.0, 0.0)
    W2 = _gemm_any(_new(m, n - m, A), None, T, W, True, False, 1.0, 0.0)
    R = torch.empty((m, n), dtype=torch.float64, device=dev)
    _copy_into(R[:, :m], R1)
    _gemm_any(R[:, m:], A2, V, W2, False, False, -1.0, 1.0)
    return V, T, R


def qr_factor_triangular(x0, x1, **kwargs):
    """QR of [triu(x0); triu(x1)] for two n x n factors (npw_tpqrt_f64).  The stacked Householder QR has exactly
    dtpqrt's reflectors (column j touches row j of the top block and rows 0..j of the bottom block only), so the
    structured LAPACK routine and the general panel kernel agree; the structured flop saving (about 2x) is not
    exploited yet."""
    _check_tile(x0, "x0")
    _check_tile(x1, "x1")
    n = x0.shape[1]
    if x0.shape[0] != n or tuple(x1.shape) != (n, n):
        raise ValueError(f"qr_factor_triangular expects two square tiles of equal size, got {tuple(x0.shape)} {tuple(x1.shape)}")
    lib = _capi.load()
    dev = x0.device
    m0, ld0, tr0 = _mat(x0, "x0")
    m1, ld1, tr1 = _mat(x1, "x1")
    if tr0:
        m0, ld0 = transpose(x0.T), n       # row-major copy (the entry point takes row-major factors)
    if tr1:
        m1, ld1 = transpose(x1.T), n
    V2 = torch.empty((n, n), dtype=torch.float64, device=dev)
    T = torch.empty((n, n), dtype=torch.float64, device=dev)
    R = torch.empty((n, n), dtype=torch.float64, device=dev)
    work = torch.empty(max(1, lib.npw_tpqrt_work_bytes(n) // 8), dtype=torch.float64, device=dev)
    rc = lib.npw_tpqrt_f64(V2.data_ptr(), max(1, n), T.data_ptr(), max(1, n), R.data_ptr(), max(1, n), m0.data_ptr(), ld0,
                           m1.data_ptr(), ld1, n, work.data_ptr(), _stream())
    _capi.check(rc, "npw_tpqrt_f64")
    if _SEMANTICS == "householder":
        return V2, T, R
    V = add_diag(_fill2d(torch.empty((n, n), dtype=torch.float64, device=x0.device), 0), 1.0)
    if n > _TPQRT_NB:
        Tb = _fill2d(torch.empty((n, n), dtype=torch.float64, device=x0.device), 0)
        for k0 in range(0, n, _TPQRT_NB):
            w = min(_TPQRT_NB, n - k0)
            _copy_into(Tb[0:w, k0:k0 + w], T[k0:k0 + w, k0:k0 + w])
        T = Tb
    return V, T, R


def lq_factor(*blocks, **kwargs):
    """fast_qr(np.hstack(blocks).T) transposed back: (v.T, t.T, r.T)."""
    if not blocks:
        raise TypeError("lq_factor expects at least one tile")
    for i, b in enumerate(blocks):
        _check_tile(b, f"blocks[{i}]")
    if len(blocks) == 2:
        assert blocks[0].shape[0] == blocks[1].shape[0]
    v, t, r = qr_factor(*[b.T for b in blocks])
    return transpose(v), transpose(t), transpose(r)


def _new(rows, cols, like):
    return torch.empty((rows, cols), dtype=torch.float64, device=like.device)


def qr_leaf(V, T, S0):
    for name, x in (("V", V), ("T", T), ("S0", S0)):
        _check_tile(x, name)
    m, c = S0.shape
    if V.shape[0] != m:
        raise ValueError(f"shapes {tuple(V.T.shape)} and {tuple(S0.shape)} not aligned")
    n = V.shape[1]
    if _SEMANTICS == "reference":
        if n != m:
            raise ValueError(f"operands could not be broadcast together with shapes {tuple(S0.shape)} ({n},{c})")
        return _gemm_any(_new(m, c, S0), S0, V, S0, True, False, -1.0, 1.0)       # S0 - V^T S0
    W = _gemm_any(_new(n, c, S0), None, V, S0, True, False, 1.0, 0.0)            # V^T S0
    W2 = _gemm_any(_new(n, c, S0), None, T, W, True, False, 1.0, 0.0)            # T^T (V^T S0)
    return _gemm_any(_new(m, c, S0), S0, V, W2, False, False, -1.0, 1.0)         # S0 - V T^T V^T S0


def lq_leaf(V, T, S0):
    for name, x in (("V", V), ("T", T), ("S0", S0)):
        _check_tile(x, name)
    c, m = S0.shape
    n = V.shape[0]
    W = _gemm_any(_new(c, n, S0), None, S0, V, False, True, 1.0, 0.0)            # S0 V^T
    W2 = _gemm_any(_new(c, n, S0), None, W, T, False, True, 1.0, 0.0)            # (S0 V^T) T^T
    return _gemm_any(_new(c, m, S0), S0, W2, V, False, False, -1.0, 1.0)         # S0 - S0 V^T T^T V


def qr_trailing_update(V, T, S0, S1):
    if S1 is None:
        z = _fill2d(torch.empty_like(S0, memory_format=torch.contiguous_format), 0)
        return qr_leaf(V, T, S0), z
    for name, x in (("V", V), ("T", T), ("S0", S0), ("S1", S1)):
        _check_tile(x, name)
    V = V[-S0.shape[0]:]
    n, c = S0.shape
    tmp = _gemm_any(_new(n, c, S0), S0, V, S1, True, False, 1.0, 1.0)            # S0 + V^T S1
    nW = _gemm_any(_new(T.shape[1], c, S0), None, T, tmp, True, False, -1.0, 0.0)  # -W = -T^T (S0 + V^T S1)
    S01 = add_matrices(S0, nW)                                                    # S0 - W
    S11 = _gemm_any(_new(S1.shape[0], c, S0), S1, V, nW, False, False, 1.0, 1.0)  # S1 - V W
    return S01, S11


def lq_trailing_update(V, T, S0, S1=None):
    if S1 is None:
        z = _fill2d(torch.empty_like(S0, memory_format=torch.contiguous_format), 0)
        return lq_leaf(V, T, S0), z
    for name, x in (("V", V), ("T", T), ("S0", S0), ("S1", S1)):
        _check_tile(x, name)
    V = V[:, -S0.shape[0]:]
    c, n = S0.shape
    tmp = _gemm_any(_new(c, n, S0), S0, S1, V, False, True, 1.0, 1.0)            # S0 + S1 V^T
    nW = _gemm_any(_new(c, T.shape[0], S0), None, tmp, T, False, True, -1.0, 0.0)  # -W = -(S0 + S1 V^T) T^T
    S01 = add_matrices(S0, nW)
    S11 = _gemm_any(_new(c, S1.shape[1], S0), S1, nW, V, False, False, 1.0, 1.0)  # S1 - W V
    assert S0.shape == S01.shape
    assert S1.shape == S11.shape
    return S01, S11
