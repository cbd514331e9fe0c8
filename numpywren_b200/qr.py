"""kernels.qr_factor on B200: compact-WY QR of the vertical stack of the argument tiles.

Reference: kernels.qr_factor (kernels.py:127-130) = fast_qr(np.vstack(blocks)) (kernels.py:86-105, LAPACK dgeqrt3).
Returns (V, T, R) as row-major CUDA tensors: V m x n unit lower trapezoidal (explicit ones / zeros), T n x n upper,
R n x n upper, Q = I - V T V^T.
"""
from __future__ import annotations

import torch

from . import _capi
from .kernels import _check_tile, _mat, _stream


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
    if n > m:
        raise _capi.NpwError("qr_factor: wide inputs (n > m) take the reference's slow_qr path, which is off the hot path")
    lib = _capi.load()
    dev = blocks[0].device
    V = torch.empty((m, n), dtype=torch.float64, device=dev)
    # np.vstack: the stacked copy becomes the working matrix that V overwrites
    r0 = 0
    for b in blocks:
        bm, ld, tr = _mat(b, "block")
        dst = V[r0:r0 + b.shape[0]]
        if tr:
            rc = lib.npw_copy2d_f64(dst.data_ptr(), max(1, n), bm.data_ptr(), ld, b.shape[1], b.shape[0], 1, _stream())
        else:
            rc = lib.npw_copy2d_f64(dst.data_ptr(), max(1, n), bm.data_ptr(), ld, b.shape[0], b.shape[1], 0, _stream())
        _capi.check(rc, "npw_copy2d_f64")
        r0 += b.shape[0]
    T = torch.empty((n, n), dtype=torch.float64, device=dev)
    R = torch.empty((n, n), dtype=torch.float64, device=dev)
    work = torch.empty(max(1, lib.npw_geqrt_work_bytes(m, n) // 8), dtype=torch.float64, device=dev)
    rc = lib.npw_geqrt_f64(V.data_ptr(), max(1, n), T.data_ptr(), max(1, n), R.data_ptr(), max(1, n), V.data_ptr(), max(1, n),
                           m, n, work.data_ptr(), _stream())
    _capi.check(rc, "npw_geqrt_f64")
    return V, T, R
