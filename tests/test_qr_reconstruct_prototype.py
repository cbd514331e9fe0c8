"""tools/qr_reconstruct_prototype.py (design study for the TSQR-leaf panel): TSQR + Householder
reconstruction must reproduce LAPACK's compact-WY factors, otherwise it cannot replace the column-by-column panel."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import qr_reconstruct_prototype as qp  # noqa: E402


@pytest.mark.parametrize("m,n,chunk", [(512, 32, 64), (2048, 32, 100), (1000, 17, 128), (4096, 64, 444)])
def test_reconstruction_equals_lapack_geqrt(m, n, chunk):
    A = np.random.RandomState(m + n).randn(m, n)
    Q, R = qp.tsqr_explicit_q(A, chunk)
    Y, T, Rh = qp.householder_from_q(Q, R)
    v, t, r = qp.lapack(A)
    assert np.abs(Y - v).max() < 1e-13 and np.abs(T - t).max() < 1e-13 and np.abs(Rh - r).max() < 1e-13 * np.abs(r).max()
    # and it is a Householder representation of A in its own right
    Qh = np.eye(m) - Y @ T @ Y.T
    assert np.abs(Qh.T @ Qh - np.eye(m)).max() < 1e-12
    assert np.abs(Qh[:, :n] @ Rh - A).max() < 1e-12 * np.abs(A).max() * n
