"""The C-ABI boundary: libnpw_b200.so loads without a GPU, exports every symbol include/npw_b200.h declares,
validates arguments before touching CUDA, and the Python side refuses to compute on the CPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from numpywren_b200 import _capi, kernels

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "npw_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(npw_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built_in_tree():
    assert os.path.exists(_capi.lib_path()), "run `python -c 'import __graft_entry__ as g; g.build()'`"
    assert _capi.lib_path().startswith(ROOT)


def test_every_declared_symbol_is_exported_and_bound():
    lib = _capi.load()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_capi.EXPORTED_SYMBOLS) == names


def test_identity_and_sizes():
    lib = _capi.load()
    assert lib.npw_version() >= 100
    assert lib.npw_build_arch() == b"sm_100a"
    assert lib.npw_invdiag_bytes(4096) == 32 * 128 * 128 * 8
    assert lib.npw_invdiag_bytes(130) == 2 * 128 * 128 * 8
    assert lib.npw_potrf_work_bytes(4096) == 4096 * 128 * 8 + 32 * 128 * 128 * 8
    assert lib.npw_trsm_work_bytes(4096, 4096) == 16          # the solve works in place: the argument is kept for ABI stability
    assert lib.npw_trsm_work_bytes(0, 10) == 0


def test_argument_validation_happens_before_cuda():
    lib = _capi.load()
    buf = (ctypes.c_double * 16)()
    p = ctypes.addressof(buf)
    assert lib.npw_syrk_f64(0, 4, p, 4, p, 4, p, 4, 4, 4, 4, None) == -1
    assert lib.npw_syrk_f64(p, 2, p, 4, p, 4, p, 4, 4, 4, 4, None) == -2     # ldc < n
    assert lib.npw_syrk_f64(p, 4, p, 4, p, 4, p, 4, 4, 4, -1, None) == -11
    assert lib.npw_gemm_f64(p, 4, 0, 0, p, 2, 0, p, 4, 0, 4, 4, 4, 1.0, 0.0, None) == -6   # lda < k
    assert lib.npw_trsm_rlt_f64(p, 4, p, 2, p, 4, 4, 4, None, None, None) == -4             # ldl < n
    assert lib.npw_potrf_l_f64(p, 4, p, 4, 4, None, None, None, None) == -6                  # no info pointer
    assert lib.npw_addn_f64(p, None, 1, 4, None) == -2
    arr = (ctypes.c_void_p * 1)(p)
    assert lib.npw_addn_f64(p, arr, 9, 4, None) == -3
    assert lib.npw_fill2d_f64(p, 4, 4, 4, 7, 0.0, None) == -5
    assert lib.npw_copy2d_f64(p, 2, p, 4, 4, 4, 0, None) == -2
    # empty problems are no-ops that never reach the device
    assert lib.npw_syrk_f64(p, 4, p, 4, p, 4, p, 4, 0, 4, 4, None) == 0
    assert lib.npw_mul_f64(p, p, p, 0, None) == 0


def test_check_turns_status_into_exception():
    with pytest.raises(_capi.NpwError, match="bad argument #3"):
        _capi.check(-3, "demo")
    _capi.check(0, "demo")


def test_no_cpu_fallback():
    a = torch.zeros(4, 4, dtype=torch.float64)
    with pytest.raises(_capi.NpwError, match="no CPU fallback"):
        kernels.syrk(a, a, a)
    with pytest.raises(_capi.NpwError):
        kernels.chol(a)
    with pytest.raises(_capi.NpwError):
        kernels.gemm(a, a)
    with pytest.raises(TypeError):
        kernels.syrk(np.zeros((4, 4)), a, a)
    for fn, args in ((kernels.qr_factor, (a,)), (kernels.qr_factor_triangular, (a, a)), (kernels.lq_factor, (a,)),
                     (kernels.qr_leaf, (a, a, a)), (kernels.lq_leaf, (a, a, a)), (kernels.qr_trailing_update, (a, a, a, a)),
                     (kernels.lq_trailing_update, (a, a, a, a))):
        with pytest.raises(_capi.NpwError, match="no CPU fallback"):
            fn(*args)


def test_flop_models_match_reference_formulas():
    # kernels.py:217-221, 228-229, 246-249, 259-263, 137-141
    s = torch.zeros(8, 6, dtype=torch.float64)
    x = torch.zeros(8, 5, dtype=torch.float64)
    y = torch.zeros(6, 5, dtype=torch.float64)
    assert kernels.syrk.flops(s, x, y) == 2 * 8 * 5 * 5 + 8 * 5
    assert kernels.chol.flops(torch.zeros(9, 9)) == 9 ** 3 / 3
    assert kernels.gemm.flops(torch.zeros(3, 4), torch.zeros(4, 5)) == 2 * 3 * 4 * 5
    assert kernels.trsm.flops(torch.zeros(4, 4), torch.zeros(7, 4)) == 4 * 4 * 4
    assert kernels.qr_factor.flops(torch.zeros(10, 4), torch.zeros(6, 4)) == 2 * 16 * 16 - 2 * 64 / 3


def test_spinning_panel_kernel_leaves_register_headroom():
    """qr_panel_reg_kernel spins on packets from its sibling CTAs, so every CTA must become resident — also on an SM where a
    tiny kernel that waits for another GPU is parked (the tile exchange's wait_signal).  With 255 registers x 256 threads it
    filled the register file and the 8-GPU TSQR deadlocked (profiles/r02i_call13_failures.txt): keep >= 4096 registers free."""
    import re
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    lib_path = os.path.join(ROOT, "numpywren_b200", "lib", "libnpw_b200.so")
    out = subprocess.run(["cuobjdump", "--dump-resource-usage", lib_path], capture_output=True, text=True).stdout
    m = re.search(r"Function [^\n]*qr_panel_reg_kernel[^\n]*:\s*\n\s*REG:(\d+)", out)
    assert m, "qr_panel_reg_kernel not found in the library"
    regs = int(m.group(1))
    assert regs * 256 <= 65536 - 4096, regs
