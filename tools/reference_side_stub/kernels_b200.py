"""The reference-side binding of INTEGRATION.md §2, as a file: what a numpywren maintainer would add as
`numpywren/kernels_b200.py` to run the tile arithmetic of `kernels.syrk / trsm / chol` on a B200 through the C-ABI of
libnpw_b200.so — NumPy arrays in and out, ctypes + the CUDA runtime for device memory, no torch, nothing from numpywren_b200.

    # numpywren/frontend.py, after line 12 (`from numpywren.kernels import *`):
    from numpywren.kernels_b200 import syrk, trsm, chol

tests/test_reference_stub_gpu.py drives exactly this module from the program logic of algs.CHOLESKY."""
import ctypes
import os

import numpy as np

try:
    from cuda.bindings import runtime as cudart
except ImportError:                                  # older cuda-python
    from cuda import cudart

_lib = ctypes.CDLL(os.environ.get("NPW_B200_LIB", "libnpw_b200.so"))
_i64, _vp = ctypes.c_int64, ctypes.c_void_p
_lib.npw_syrk_f64.argtypes = [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _vp]
_lib.npw_trsm_rlt_f64.argtypes = [_vp, _i64, _vp, _i64, _vp, _i64, _i64, _i64, _vp, _vp, _vp]
_lib.npw_potrf_l_f64.argtypes = [_vp, _i64, _vp, _i64, _i64, _vp, _vp, _vp, _vp]
_lib.npw_trsm_work_bytes.restype = _lib.npw_potrf_work_bytes.restype = ctypes.c_size_t
_lib.npw_trsm_work_bytes.argtypes = [_i64, _i64]
_lib.npw_potrf_work_bytes.argtypes = [_i64]
_lib.npw_geqrt_f64.argtypes = [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _i64, _vp, _vp]
_lib.npw_geqrt_work_bytes.restype = ctypes.c_size_t
_lib.npw_geqrt_work_bytes.argtypes = [_i64, _i64]
_H2D, _D2H = cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost


def _malloc(nbytes):
    err, p = cudart.cudaMalloc(max(int(nbytes), 16))
    assert int(err) == 0, err
    return int(p)


def _to_dev(a):                              # C-ordered fp64 ndarray -> device pointer
    a = np.ascontiguousarray(a, dtype=np.float64)
    p = _malloc(a.nbytes)
    (err,) = cudart.cudaMemcpy(p, a.ctypes.data, a.nbytes, _H2D)
    assert int(err) == 0, err
    return p, a.shape


def _to_host(p, shape):
    out = np.empty(shape)
    (err,) = cudart.cudaMemcpy(out.ctypes.data, p, out.nbytes, _D2H)    # synchronises with the kernels on the default stream
    assert int(err) == 0, err
    cudart.cudaFree(p)
    return out


def syrk(s, x, y, *args, **kwargs):          # replaces kernels.syrk (kernels.py:212-215)
    ps, (m, n) = _to_dev(s)
    px, (_, k) = _to_dev(x)
    py, _ = _to_dev(y)
    rc = _lib.npw_syrk_f64(ps, n, ps, n, px, k, py, k, m, n, k, None)
    assert rc == 0, rc
    out = _to_host(ps, (m, n))
    cudart.cudaFree(px)
    cudart.cudaFree(py)
    return out


def trsm(x, y, lower=False, right=True, *args, **kwargs):   # kernels.trsm (kernels.py:254-257), the DSL's call form
    pl, (n, _) = _to_dev(x)
    pb, (m, _) = _to_dev(y)
    w = _malloc(_lib.npw_trsm_work_bytes(m, n))
    rc = _lib.npw_trsm_rlt_f64(pb, n, pl, n, pb, n, m, n, None, w, None)
    assert rc == 0, rc
    out = _to_host(pb, (m, n))
    cudart.cudaFree(w)
    cudart.cudaFree(pl)
    return out


def chol(x, *args, **kwargs):                # kernels.chol (kernels.py:225-226)
    pa, (n, _) = _to_dev(x)
    w = _malloc(_lib.npw_potrf_work_bytes(n))
    info = _malloc(4)
    rc = _lib.npw_potrf_l_f64(pa, n, pa, n, n, info, None, w, None)
    assert rc == 0, rc
    code = np.zeros(1, np.int32)
    cudart.cudaMemcpy(code.ctypes.data, info, 4, _D2H)
    cudart.cudaFree(w)
    cudart.cudaFree(info)
    if code[0]:
        cudart.cudaFree(pa)
        raise np.linalg.LinAlgError("Matrix is not positive definite")
    return _to_host(pa, (n, n))


def qr_factor(*blocks, **kwargs):            # kernels.qr_factor (kernels.py:127-130): fast_qr(np.vstack(blocks)) -> (V, T, R)
    a = np.vstack(blocks)
    m, n = a.shape
    assert m >= n, "the wide form (slow_qr) is sequenced by numpywren_b200/qr.py from the same entry points"
    pv, _ = _to_dev(a)
    pt, pr = _malloc(n * n * 8), _malloc(n * n * 8)
    w = _malloc(_lib.npw_geqrt_work_bytes(m, n))
    rc = _lib.npw_geqrt_f64(pv, n, pt, n, pr, n, pv, n, m, n, w, None)
    assert rc == 0, rc
    V, T, R = _to_host(pv, (m, n)), _to_host(pt, (n, n)), _to_host(pr, (n, n))
    cudart.cudaFree(w)
    return V, T, R
