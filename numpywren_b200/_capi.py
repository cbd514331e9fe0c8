"""ctypes binding of libnpw_b200.so (include/npw_b200.h).

This is the ONLY way the Python host side reaches the CUDA kernels: raw device
pointers + leading dimensions + a cudaStream_t, exactly the C-ABI a cgo/JNI/ctypes
binding on the reference side would use.  There is no CPU
fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libnpw_b200.so")

NPW_ERR_CUDA = -1000
NPW_ERR_UNSUPPORTED = -1001
DIAG_NB = 128


class NpwError(RuntimeError):
    """A libnpw_b200 call returned a non-zero status."""


_lib = None

# name -> (restype, argtypes); mirrors include/npw_b200.h one to one.
_SIGNATURES = {
    "npw_version": (c_int, []),
    "npw_last_error": (c_char_p, []),
    "npw_build_arch": (c_char_p, []),
    "npw_launch_count": (c_uint64, []),
    "npw_syrk_f64": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                             c_int64, c_int64, c_int64, c_void_p]),
    "npw_syrk_lower_f64": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                                   c_int64, c_int64, c_int64, c_void_p]),
    "npw_gemm_f64": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_int64, c_int,
                             c_int64, c_int64, c_int64, c_double, c_double, c_void_p]),
    "npw_trsm_work_bytes": (c_size_t, [c_int64, c_int64]),
    "npw_trsm_rlt_f64": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64,
                                 c_void_p, c_void_p, c_void_p]),
    "npw_invdiag_bytes": (c_size_t, [c_int64]),
    "npw_trtri_diag_f64": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "npw_potrf_work_bytes": (c_size_t, [c_int64]),
    "npw_potrf_l_f64": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "npw_addn_f64": (c_int, [c_void_p, ctypes.POINTER(c_void_p), c_int, c_int64, c_void_p]),
    "npw_mul_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "npw_copy2d_f64": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p]),
    "npw_add_diag_f64": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_double, c_void_p]),
    "npw_fill2d_f64": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int, c_double, c_void_p]),
    "npw_geqrt_work_bytes": (c_size_t, [c_int64, c_int64]),
    "npw_geqrt_f64": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                              c_int64, c_int64, c_void_p, c_void_p]),
    "npw_tpqrt_work_bytes": (c_size_t, [c_int64]),
    "npw_tpqrt_f64": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                              c_int64, c_void_p, c_void_p]),
    "npw_i8_digits_bytes": (c_size_t, [c_int64, c_int64, c_int]),
    "npw_split_i8_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p]),
    "npw_syrk_i8emu_f64": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_int64, c_int64, c_int64, c_int, c_int, c_void_p]),
    "npw_fp64_pipe_probe_bytes": (c_size_t, [c_int]),
    "npw_fp64_pipe_probe": (c_int, [c_void_p, c_int, c_int, ctypes.POINTER(c_double), c_void_p]),
    "npw_fill_random_f64": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_uint64, c_int64, c_int64, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib_path() -> str:
    return _LIB_PATH


def load():
    """Load libnpw_b200.so (once) and attach prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise NpwError(
            f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C numpywren_b200/csrc` (there is no CPU fallback)")
    lib = ctypes.CDLL(_LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    lib = load()
    msg = lib.npw_last_error()
    msg = msg.decode() if msg else ""
    if rc == NPW_ERR_CUDA:
        raise NpwError(f"{what}: CUDA error: {msg}")
    if rc == NPW_ERR_UNSUPPORTED:
        raise NpwError(f"{what}: unsupported: {msg}")
    raise NpwError(f"{what}: bad argument #{-rc} ({msg})")


def launch_count() -> int:
    return int(load().npw_launch_count())
