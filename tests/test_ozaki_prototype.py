"""tools/ozaki_prototype.py (design study for the int8-tensor-core fp64 emulation): the accuracy figures
quoted there must keep holding, and the int32 exactness argument must not silently break."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ozaki_prototype as oz  # noqa: E402


def test_digits_reconstruct_the_operand():
    rs = np.random.RandomState(0)
    A = rs.randn(37, 129) * np.exp(rs.uniform(-20, 20, size=37))[:, None]
    A[3] = 0.0                                           # an all-zero row must not produce NaNs
    A[5, 0] = 2.0 ** 10                                  # exact power of two as the row maximum
    d, e, w = oz.split_rows(A, 8)
    assert d.dtype == np.int8 and np.abs(d.astype(int)).max() <= 64
    rec = sum(d[p].astype(np.float64) * 2.0 ** -w[p] for p in range(8)) * np.exp2(e)[:, None]
    scale = np.abs(A).max(axis=1, keepdims=True) + (np.abs(A).max(axis=1, keepdims=True) == 0)
    assert (np.abs(rec - A) / scale).max() < 2.0 ** -54


@pytest.mark.parametrize("s,tol", [(4, 2e-7), (5, 2e-9), (6, 2e-11), (7, 2e-13), (8, 2e-15)])
def test_emulated_product_accuracy(s, tol):
    rs = np.random.RandomState(s)
    X, Y = rs.randn(96, 512), rs.randn(80, 512)
    X *= np.exp(rs.uniform(-14, 14, size=96))[:, None]
    C, nprod = oz.ozaki_gemm_nt(X, Y, s)
    assert nprod == s * (s + 1) // 2
    ref = np.asarray(X.astype(np.longdouble) @ Y.T.astype(np.longdouble), dtype=np.float64)
    assert oz.rel(C, ref) < tol


def test_int32_group_sums_cannot_overflow_at_the_benchmark_tile():
    # worst case: every digit is +-64, k = 4096, s = 8 digits -> the largest group holds 8 pairs
    k, s = 4096, 8
    assert k * 64 * 64 * s < 2 ** 31
