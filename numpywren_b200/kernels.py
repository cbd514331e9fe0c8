"""Tile kernels of the LambdaPACK hot path on B200.

Same names, positional arguments and return conventions as the reference's
``numpywren/kernels.py`` (looked up *by name* from the DSL, frontend.py:343), but every
function takes and returns ``torch`` CUDA float64 tensors and runs a hand-written
sm_100a kernel from ``libnpw_b200.so`` on the current CUDA stream.  Nothing here
computes on the CPU; without the shared library every call raises.

Deviations from the reference that do not change results:
  * ``syrk``/``trsm`` do not scan their inputs with ``np.allclose(.,0)`` (kernels.py:213,255).
    The short-circuit returns ``s`` (resp. zeros) when an operand is ~0, which equals the
    computed value to one rounding of zero; skipping two O(b^2) scans keeps the call async.
  * ``trsm`` returns a C-ordered tile (the reference returns the F-ordered BLAS buffer).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _capi

__all__ = [
    "add_matrices", "syrk", "chol", "trsm", "gemm", "mul", "identity", "qr_factor",
    "qr_factor_triangular", "qr_leaf", "qr_trailing_update", "lq_factor", "lq_leaf", "lq_trailing_update",
    "gemm_acc", "chol_async", "trsm_with_inverse", "transpose", "add_diag", "fill_random",
]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _check_tile(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor on a CUDA device, got {type(t).__name__}")
    if not t.is_cuda:
        raise _capi.NpwError(f"{name}: tensor is on {t.device}; the B200 kernels have no CPU fallback")
    if t.dtype != torch.float64:
        raise TypeError(f"{name}: expected float64, got {t.dtype}")
    if t.dim() != 2:
        raise ValueError(f"{name}: expected a 2-D tile, got shape {tuple(t.shape)}")


def _mat(t, name):
    """-> (tensor keeping the memory alive, leading dimension, stored_transposed)."""
    _check_tile(t, name)
    r, c = t.shape
    if t.stride(1) == 1 and t.stride(0) >= max(1, c):
        return t, t.stride(0), False
    if t.stride(0) == 1 and t.stride(1) >= max(1, r):
        return t, t.stride(1), True
    t = t.contiguous()
    return t, max(1, c), False


def _rowmajor(t, name):
    _check_tile(t, name)
    if t.stride(1) == 1 and t.stride(0) >= max(1, t.shape[1]):
        return t
    return _materialize(t)


def _materialize(t):
    """Row-major copy of an arbitrarily strided 2-D tile using the library's own copy kernels."""
    base, ld, tr = _mat(t, "tile")
    out = torch.empty(t.shape, dtype=torch.float64, device=t.device)
    lib = _capi.load()
    if tr:
        # memory holds t^T row-major (cols x rows); out[r, c] = stored[c, r]
        rc = lib.npw_copy2d_f64(out.data_ptr(), max(1, out.shape[1]), base.data_ptr(), ld, t.shape[1], t.shape[0], 1, _stream())
    else:
        rc = lib.npw_copy2d_f64(out.data_ptr(), max(1, out.shape[1]), base.data_ptr(), ld, t.shape[0], t.shape[1], 0, _stream())
    _capi.check(rc, "npw_copy2d_f64")
    return out


def _gemm_into(out, c0, a, b, trans_a, trans_b, alpha, beta):
    """out = alpha * op(a) @ op(b) + beta * c0, honouring stored-transposed operands without copies."""
    lib = _capi.load()
    am, lda, a_t = _mat(a, "A")
    bm, ldb, b_t = _mat(b, "B")
    ta = bool(trans_a) ^ a_t
    tb = bool(trans_b) ^ b_t
    m, k = (a.shape[1], a.shape[0]) if trans_a else (a.shape[0], a.shape[1])
    k2, n = (b.shape[1], b.shape[0]) if trans_b else (b.shape[0], b.shape[1])
    if k != k2:
        raise ValueError(f"shapes {tuple(a.shape)} and {tuple(b.shape)} not aligned: {k} (dim 1) != {k2} (dim 0)")
    if tuple(out.shape) != (m, n):
        raise ValueError(f"output shape {tuple(out.shape)} != ({m}, {n})")
    c0p, ldc0 = (0, 0)
    if c0 is not None and beta != 0.0:
        c0 = _rowmajor(c0, "C0")
        if tuple(c0.shape) != (m, n):
            raise ValueError(f"operands could not be broadcast together with shapes {tuple(c0.shape)} ({m},{n})")
        c0p, ldc0 = c0.data_ptr(), c0.stride(0)
    rc = lib.npw_gemm_f64(out.data_ptr(), out.stride(0), c0p, ldc0, am.data_ptr(), lda, int(ta), bm.data_ptr(), ldb, int(tb),
                          m, n, k, float(alpha), float(beta), _stream())
    _capi.check(rc, "npw_gemm_f64")
    return out


# ----------------------------------------------------------------------------- reference API
def add_matrices(*args, **kwargs):
    """zeros(args[0].shape) + sum(args)  — kernels.py:16-20."""
    if not args:
        raise TypeError("add_matrices expects at least one tile")
    lib = _capi.load()
    tiles = [_rowmajor(a, f"args[{i}]").contiguous() for i, a in enumerate(args)]
    shape = tiles[0].shape
    for t in tiles[1:]:
        if t.shape != shape:
            raise ValueError(f"operands could not be broadcast together with shapes {tuple(shape)} {tuple(t.shape)}")
    out = torch.empty(shape, dtype=torch.float64, device=tiles[0].device)
    nelem = out.numel()
    # the kernel adds up to 8 operands per pass, in argument order like the reference loop
    chunk, rest = tiles[:8], tiles[8:]
    while True:
        arr = (ctypes.c_void_p * len(chunk))(*[t.data_ptr() for t in chunk])
        rc = lib.npw_addn_f64(out.data_ptr(), arr, len(chunk), nelem, _stream())
        _capi.check(rc, "npw_addn_f64")
        if not rest:
            break
        chunk, rest = [out] + rest[:7], rest[7:]
    return out


def _out_tile(out, shape, device, name):
    """Validate a caller-provided output tile (row-major, right shape) or allocate one."""
    if out is None:
        return torch.empty(shape, dtype=torch.float64, device=device)
    _check_tile(out, name)
    if tuple(out.shape) != tuple(shape) or out.stride(1) != 1 or out.stride(0) < max(1, shape[1]):
        raise ValueError(f"{name}: expected a row-major {tuple(shape)} tile, got {tuple(out.shape)} strides {out.stride()}")
    return out


def syrk(s, x, y, *args, out=None, lower=False, **kwargs):
    """s - x.dot(y.T)  — kernels.py:212-215.  ``out`` (scheduler use) may be ``s`` itself: in-place update.
    ``lower=True`` (scheduler use, diagonal tiles) updates only the CTA tiles touching the lower triangle."""
    _check_tile(s, "s")
    _check_tile(x, "x")
    _check_tile(y, "y")
    out = _out_tile(out, (x.shape[0], y.shape[0]), s.device, "out")
    if tuple(s.shape) != tuple(out.shape):
        raise ValueError(f"operands could not be broadcast together with shapes {tuple(s.shape)} {tuple(out.shape)}")
    sm = _rowmajor(s, "s")
    xm, ldx, x_t = _mat(x, "x")
    ym, ldy, y_t = _mat(y, "y")
    if not x_t and not y_t:
        if x.shape[1] != y.shape[1]:
            raise ValueError(f"shapes {tuple(x.shape)} and {tuple(y.T.shape)} not aligned")
        fn = _capi.load().npw_syrk_lower_f64 if lower else _capi.load().npw_syrk_f64
        rc = fn(out.data_ptr(), out.stride(0), sm.data_ptr(), sm.stride(0), xm.data_ptr(), ldx,
                ym.data_ptr(), ldy, x.shape[0], y.shape[0], x.shape[1], _stream())
        _capi.check(rc, "npw_syrk_f64")
        return out
    return _gemm_into(out, sm, x, y, False, True, -1.0, 1.0)


def _syrk_flops(s, x, y):
    m, n = x.shape
    z = y.shape[1]
    return 2 * m * n * z + m * z


syrk.flops = _syrk_flops


def chol_async(x, want_inverse=True, out=None):
    """Enqueue the tile Cholesky; returns (L, info[int32 device scalar], invdiag or None).

    No host synchronisation: ``info`` holds LAPACK's INFO once the stream reaches it.
    ``invdiag`` (inverted 128x128 diagonal blocks of L) feeds ``trsm_with_inverse``.
    """
    x = _rowmajor(x, "x")
    n = x.shape[0]
    if x.shape[1] != n:
        raise np.linalg.LinAlgError("Last 2 dimensions of the array must be square")
    lib = _capi.load()
    dev = x.device
    L = _out_tile(out, (n, n), dev, "out")
    info = torch.empty((), dtype=torch.int32, device=dev)
    work = torch.empty(max(1, lib.npw_potrf_work_bytes(n) // 8), dtype=torch.float64, device=dev)
    inv = None
    if want_inverse:
        inv = torch.empty(max(1, lib.npw_invdiag_bytes(n) // 8), dtype=torch.float64, device=dev)
    rc = lib.npw_potrf_l_f64(L.data_ptr(), max(1, L.stride(0)), x.data_ptr(), max(1, x.stride(0)), n, info.data_ptr(),
                             inv.data_ptr() if inv is not None else 0, work.data_ptr(), _stream())
    _capi.check(rc, "npw_potrf_l_f64")
    return L, info, inv


def chol(x, *args, **kwargs):
    """np.linalg.cholesky(x)  — kernels.py:225-226 (raises LinAlgError when x is not SPD)."""
    L, info, _ = chol_async(x, want_inverse=False)
    code = int(info.item())
    if code != 0:
        raise np.linalg.LinAlgError("Matrix is not positive definite")
    return L


def _chol_flops(x):
    return (x.shape[0] ** 3) / 3


chol.flops = _chol_flops


def trsm_with_inverse(x, y, invdiag=None, out=None):
    """y @ inv(x).T for lower-triangular x, optionally reusing chol_async's inverted diagonal blocks."""
    _check_tile(x, "x")
    y = _rowmajor(y, "y")
    n = x.shape[0]
    if x.shape[1] != n:
        raise ValueError(f"trsm: triangular factor must be square, got {tuple(x.shape)}")
    m = y.shape[0]
    if y.shape[1] != n:
        raise ValueError(f"trsm: shapes {tuple(x.shape)} and {tuple(y.shape)} are incompatible")
    xm = _rowmajor(x, "x")
    lib = _capi.load()
    out = _out_tile(out, (m, n), y.device, "out")
    work = torch.empty(max(1, lib.npw_trsm_work_bytes(m, n) // 8), dtype=torch.float64, device=y.device)
    rc = lib.npw_trsm_rlt_f64(out.data_ptr(), max(1, out.stride(0)), xm.data_ptr(), max(1, xm.stride(0)), y.data_ptr(),
                              max(1, y.stride(0)), m, n, invdiag.data_ptr() if invdiag is not None else 0, work.data_ptr(),
                              _stream())
    _capi.check(rc, "npw_trsm_rlt_f64")
    return out


def trsm(x, y, lower=False, right=True, *args, **kwargs):
    """scipy.linalg.blas.dtrsm(1.0, x.T, y, lower=lower, side=int(right))  — kernels.py:254-257.

    The DSL always calls it with the defaults (frontend.py:346 drops keyword arguments): ``y @ inv(tril(x)).T``, which is
    the one triangular solve libnpw_b200 implements (npw_trsm_rlt_f64).  The other three flag combinations are mapped
    onto it by data movement only — transposes, and index reversal J (J U J is lower triangular when U is upper):
      right, lower=True : y inv(triu(x)).T       = ((y J) inv(J triu(x) J).T) J
      left,  lower=True : inv(triu(x)).T y       = (y.T inv(triu(x)).T... ) see below
      left,  lower=False: inv(tril(x)).T y       = (y.T inv(tril(x)))^T  = ((y.T J) inv(J tril(x).T J).T J).T
    """
    _check_tile(x, "x")
    _check_tile(y, "y")
    if right and not lower:
        return trsm_with_inverse(x, y, None)
    if right and lower:
        # A = tril(x.T) = triu(x).T : solve X A = y  ->  X = y inv(U).T, U = triu(x); L' = J U J is lower triangular
        Lp = torch.flip(x, (0, 1)).contiguous()
        return torch.flip(trsm_with_inverse(Lp, torch.flip(y, (1,)).contiguous(), None), (1,)).contiguous()
    if lower:
        # left, A = tril(x.T) = triu(x).T =: L'' (lower): X = inv(L'') y  ->  X.T = y.T inv(L'').T  with L'' = triu(x).T
        return transpose(trsm_with_inverse(transpose(x), transpose(y), None))
    # left, A = triu(x.T) = tril(x).T =: U' (upper): X = inv(U') y  ->  X.T = y.T inv(U').T ; J U' J is lower triangular
    Lp = torch.flip(transpose(x), (0, 1)).contiguous()
    Xt = torch.flip(trsm_with_inverse(Lp, torch.flip(transpose(y), (1,)).contiguous(), None), (1,))
    return transpose(Xt.contiguous())


def _trsm_flops(x, y):
    if len(y.shape) == 0:
        return x.shape[0] * x.shape[1]
    return x.shape[0] * x.shape[1] * y.shape[1]


trsm.flops = _trsm_flops


def mul(x, y, *args, **kwargs):
    """x * y  — kernels.py:233-234."""
    x = _rowmajor(x, "x").contiguous()
    y = _rowmajor(y, "y").contiguous()
    if x.shape != y.shape:
        raise ValueError(f"operands could not be broadcast together with shapes {tuple(x.shape)} {tuple(y.shape)}")
    out = torch.empty_like(x)
    rc = _capi.load().npw_mul_f64(out.data_ptr(), x.data_ptr(), y.data_ptr(), out.numel(), _stream())
    _capi.check(rc, "npw_mul_f64")
    return out


def identity(x, *args, **kwargs):
    """x  — kernels.py:236-237 (the tile is passed through, not copied, like the reference)."""
    return x


def gemm(A, B, *args, **kwargs):
    """op(A).dot(op(B))  — kernels.py:239-244."""
    ta = bool(kwargs.get("transpose_A", False))
    tb = bool(kwargs.get("transpose_B", False))
    _check_tile(A, "A")
    _check_tile(B, "B")
    m = A.shape[1] if ta else A.shape[0]
    n = B.shape[0] if tb else B.shape[1]
    out = torch.empty((m, n), dtype=torch.float64, device=A.device)
    # The DMMA core wants both operands K-contiguous ("NT").  op(B) that is N-contiguous is
    # re-laid-out once by the HBM-bound transpose kernel (2 x tile bytes of traffic against
    # 2mnk flops): large tiles only, small ones go through the generic kernel directly.
    bm, ldb, b_t = _mat(B, "B")
    if not (tb ^ b_t) and min(m, n) >= 256:
        Bt = transpose(B.T if tb else B)  # (n x k) row-major == op(B)^T
        return _gemm_into(out, None, A, Bt, ta, True, 1.0, 0.0)
    return _gemm_into(out, None, A, B, ta, tb, 1.0, 0.0)


def _gemm_accumulate(acc, A, B):
    """acc += A @ B in place (the serial reduction loop of binops._gemm_remote_0, binops.py:19-33)."""
    bm, ldb, b_t = _mat(B, "B")
    if not b_t and min(acc.shape) >= 256:
        return _gemm_into(acc, acc, A, transpose(B), False, True, 1.0, 1.0)
    return _gemm_into(acc, acc, A, B, False, False, 1.0, 1.0)


def gemm_acc(c, a, b, *args, out=None, **kwargs):
    """c + a.dot(b): one step of the serial reduction loop of the legacy binops GEMM
    (``XY_block += X.get_block(i, r).dot(Y.get_block(r, j))``, binops.py:19-33) as a tile kernel, so that the same
    owner-computes / K-loop schedule can run as a LambdaPACK program (algs.GEMM_ACC) on the DAG engine and across
    GPUs.  ``out`` (scheduler use) may be ``c`` itself: in-place accumulation."""
    _check_tile(c, "c")
    _check_tile(a, "a")
    _check_tile(b, "b")
    if a.shape[1] != b.shape[0]:
        raise ValueError(f"shapes {tuple(a.shape)} and {tuple(b.shape)} not aligned: {a.shape[1]} (dim 1) != {b.shape[0]} (dim 0)")
    out = _out_tile(out, (a.shape[0], b.shape[1]), c.device, "out")
    if tuple(c.shape) != tuple(out.shape):
        raise ValueError(f"operands could not be broadcast together with shapes {tuple(c.shape)} {tuple(out.shape)}")
    return _gemm_any(out, c, a, b, False, False, 1.0, 1.0)


def _gemm_acc_flops(c, a, b):
    m, n = a.shape
    return 2 * m * n * b.shape[1] + m * b.shape[1]


gemm_acc.flops = _gemm_acc_flops


def _gemm_flops(A, B):
    m, n = A.shape
    k = B.shape[1]
    return 2 * m * n * k


gemm.flops = _gemm_flops


def qr_factor(*blocks, **kwargs):
    """(V, T, R) of the compact-WY QR of vstack(blocks)  — kernels.py:127-130 → fast_qr :86-105."""
    from . import qr as _qr
    return _qr.qr_factor(*blocks, **kwargs)


def _qr_flops(*blocks):
    m = sum(b.shape[0] for b in blocks)
    n = blocks[0].shape[1]
    return 2 * m * n * n - (2 * n ** 3) / 3


qr_factor.flops = _qr_flops


def qr_factor_triangular(x0, x1, **kwargs):
    """(V, T, R) of the QR of two stacked upper-triangular factors  — kernels.py:132-134 → fast_qr_triangular :107-124."""
    from . import qr as _qr
    return _qr.qr_factor_triangular(x0, x1, **kwargs)


def lq_factor(*blocks, **kwargs):
    """(V^T, T^T, L) of the LQ of hstack(blocks)  — kernels.py:145-150."""
    from . import qr as _qr
    return _qr.lq_factor(*blocks, **kwargs)


lq_factor.flops = _qr_flops


def qr_leaf(V, T, S0, *args, **kwargs):
    """Apply a leaf reflector block to a trailing tile  — kernels.py:160-164 (see qr.py for the two semantics)."""
    from . import qr as _qr
    return _qr.qr_leaf(V, T, S0)


def lq_leaf(V, T, S0, *args, **kwargs):
    """S0 - S0 V^T T^T V  — kernels.py:154-157."""
    from . import qr as _qr
    return _qr.lq_leaf(V, T, S0)


def _qr_leaf_flops(V, T, S0):
    c0 = V.shape[0] * S0.shape[0] * S0.shape[1]
    c1 = T.shape[0] * V.shape[0] * S0.shape[1]
    c2 = V.shape[0] * T.shape[0] * T.shape[1]
    return c0 + c1 + c2 + S0.shape[0] * S0.shape[1]


qr_leaf.flops = _qr_leaf_flops
lq_leaf.flops = _qr_leaf_flops


def qr_trailing_update(V, T, S0, S1, *args, **kwargs):
    """Apply a tree-merge reflector block to a pair of trailing tiles  — kernels.py:181-188."""
    from . import qr as _qr
    return _qr.qr_trailing_update(V, T, S0, S1)


def lq_trailing_update(V, T, S0, S1=None, *args, **kwargs):
    """Row-wise mirror of qr_trailing_update  — kernels.py:199-208."""
    from . import qr as _qr
    return _qr.lq_trailing_update(V, T, S0, S1)


def _qr_trailing_flops(V, T, S0, S1):
    M, N = V.shape
    c0 = M * S1.shape[0] * S1.shape[1]
    c1 = T.shape[0] * T.shape[1] * S0.shape[1]
    return 2 * c1 + c0 + T.shape[0] * T.shape[1]


qr_trailing_update.flops = _qr_trailing_flops
lq_trailing_update.flops = _qr_trailing_flops


def _gemm_any(out, c0, A, B, trans_a, trans_b, alpha, beta):
    """out = alpha * op(A) @ op(B) + beta * c0 for any operand layout.  Large products are brought to the DMMA core's
    NT form (both operands K-contiguous) by re-laying an operand out with the HBM-bound transpose kernel
    (2 x tile bytes of traffic against 2mnk flops); small ones go straight to the generic kernel."""
    m = A.shape[1] if trans_a else A.shape[0]
    k = A.shape[0] if trans_a else A.shape[1]
    n = B.shape[0] if trans_b else B.shape[1]
    if min(m, n) >= 128 and k >= 64:
        X = A.T if trans_a else A            # op(A), m x k
        if _mat(X, "A")[2]:
            X = transpose(X.T)               # row-major copy of op(A)
        Y = B if trans_b else B.T            # op(B)^T, n x k
        if _mat(Y, "B")[2]:
            Y = transpose(Y.T)               # row-major copy of op(B)^T
        return _gemm_into(out, c0, X, Y, False, True, alpha, beta)
    return _gemm_into(out, c0, A, B, trans_a, trans_b, alpha, beta)


# ----------------------------------------------------------------------------- optional: int8-tensor-core emulation
def split_i8(x, digits=6):
    """Optional path (off by default, tests/test_i8emu_gpu.py).  fp64 tile -> (int8 digit planes [digits, rows, k],
    int32 row exponents) for ``syrk_i8emu``: x = diag(2^e) sum_p 2^-(6+7p) x_p."""
    x = _rowmajor(x, "x")
    rows, k = x.shape
    lib = _capi.load()
    d = torch.empty((digits, rows, k), dtype=torch.int8, device=x.device)
    e = torch.empty((rows,), dtype=torch.int32, device=x.device)
    rc = lib.npw_split_i8_f64(d.data_ptr(), e.data_ptr(), x.data_ptr(), max(1, x.stride(0)), rows, k, int(digits), _stream())
    _capi.check(rc, "npw_split_i8_f64")
    return d, e


def syrk_i8emu(s, xd, xe, yd, ye, out=None, lower=False):
    """Optional path.  s - x.dot(y.T) from the int8 digits of x and y (``split_i8``): the product runs on the int8 tensor
    cores (tcgen05.mma kind::i8) with exact int32 group sums, the combination in fp64."""
    _check_tile(s, "s")
    sm = _rowmajor(s, "s")
    digits, m, k = xd.shape
    n = yd.shape[1]
    if yd.shape[0] != digits or yd.shape[2] != k:
        raise ValueError(f"digit tensors do not match: {tuple(xd.shape)} vs {tuple(yd.shape)}")
    out = _out_tile(out, (m, n), s.device, "out")
    if tuple(sm.shape) != (m, n):
        raise ValueError(f"operands could not be broadcast together with shapes {tuple(s.shape)} ({m},{n})")
    rc = _capi.load().npw_syrk_i8emu_f64(out.data_ptr(), out.stride(0), sm.data_ptr(), sm.stride(0), xd.data_ptr(), xe.data_ptr(),
                                         yd.data_ptr(), ye.data_ptr(), m, n, k, int(digits), int(bool(lower)), _stream())
    _capi.check(rc, "npw_syrk_i8emu_f64")
    return out


# ----------------------------------------------------------------------------- helpers (not in the reference)
def transpose(x):
    """Row-major copy of x.T (BigMatrixView transposed reads, matrix.py:643-661)."""
    xm, ld, x_t = _mat(x, "x")
    r, c = x.shape
    out = torch.empty((c, r), dtype=torch.float64, device=x.device)
    lib = _capi.load()
    if x_t:  # memory already holds x.T row-major
        rc = lib.npw_copy2d_f64(out.data_ptr(), max(1, r), xm.data_ptr(), ld, c, r, 0, _stream())
    else:
        rc = lib.npw_copy2d_f64(out.data_ptr(), max(1, r), xm.data_ptr(), ld, r, c, 1, _stream())
    _capi.check(rc, "npw_copy2d_f64")
    return out


def add_diag(x, lambdav):
    """In-place x[i, i] += lambdav (BigMatrix.get_block's diagonal shift, matrix.py:307-309)."""
    x_rm = _rowmajor(x, "x")
    if x_rm is not x:
        raise ValueError("add_diag needs a row-major tile")
    rc = _capi.load().npw_add_diag_f64(x.data_ptr(), max(1, x.stride(0)), x.shape[0], x.shape[1], float(lambdav), _stream())
    _capi.check(rc, "npw_add_diag_f64")
    return x


def fill_random(out, seed, row0=0, col0=0):
    """Device-side synthetic fill: out[r, c] = U(-1,1) keyed by (seed, row0 + r, col0 + c)."""
    _check_tile(out, "out")
    rc = _capi.load().npw_fill_random_f64(out.data_ptr(), max(1, out.stride(0)), out.shape[0], out.shape[1], int(seed),
                                          int(row0), int(col0), _stream())
    _capi.check(rc, "npw_fill_random_f64")
    return out
