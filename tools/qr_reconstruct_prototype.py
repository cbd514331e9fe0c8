"""Numerical prototype (CPU, NumPy) for a faster QR panel: TSQR + Householder reconstruction (Ballard, Demmel, Grigori,
Jacquelin, Nguyen, Solomonik 2014).  Design study for the TSQR leaf kernel (the 65536 x 512 leaf spends 55 %
of its time in a panel kernel that needs one grid-wide reduction per COLUMN); nothing here is product code.

Idea: per 32-column panel, (1) every CTA factors its own row chunk locally (no grid sync), (2) the stacked chunk R's
(148 x 32 x 32) are factored by one CTA, (3) every CTA forms its rows of the explicit Q = Q_chunk Q_tree, (4) the
Householder representation is recovered from Q by an LU factorisation with sign choice: Q - [S; 0] = Y U,
T = -U S Y1^-T, R_h = S R.  Two or three grid-wide syncs per panel instead of 32 — and the result is the SAME (V, T, R)
LAPACK's dgeqrt returns (the Householder representation is unique), so parity with kernels.qr_factor is kept.

  python tools/qr_reconstruct_prototype.py     # prints max deviations from scipy's dgeqrt"""
import numpy as np
import scipy.linalg


def tsqr_explicit_q(A, chunk):
    """TSQR of the tall panel A: local QR per row chunk, QR of the stacked R factors, explicit Q (m x n) and R."""
    m, n = A.shape
    qs, rs = [], []
    for r0 in range(0, m, chunk):
        q, r = np.linalg.qr(A[r0:r0 + chunk])
        qs.append(q)
        rs.append(r)
    q2, R = np.linalg.qr(np.vstack(rs))
    Q = np.vstack([qs[c] @ q2[c * n:(c + 1) * n] if qs[c].shape[1] == n else qs[c] @ q2[c * n:c * n + qs[c].shape[1]]
                   for c in range(len(qs))])
    return Q, R


def householder_from_q(Q, R):
    """Q (m x n, orthonormal columns), R -> (Y, T, R_h) with I - Y T Y^T the Householder QR of Q R in LAPACK's
    convention (unit lower-trapezoidal Y, upper-triangular T, R_h = S R)."""
    m, n = Q.shape
    W = Q.copy()
    S = np.zeros(n)
    for i in range(n):                     # LU without pivoting of Q - [S; 0], the sign picked at each pivot
        S[i] = -1.0 if W[i, i] >= 0 else 1.0
        W[i, i] -= S[i]
        W[i + 1:, i] /= W[i, i]
        W[i + 1:, i + 1:] -= np.outer(W[i + 1:, i], W[i, i + 1:])
    Y = np.tril(W, -1)
    Y[np.arange(n), np.arange(n)] = 1.0
    U = np.triu(W[:n])
    T = -(U * S[None, :]) @ np.linalg.inv(Y[:n]).T
    return Y, np.triu(T), S[:, None] * R


def lapack(A):
    n = A.shape[1]
    a, t, info = scipy.linalg.lapack.dgeqrt(n, np.asfortranarray(A))
    v = np.tril(a, -1)
    v[np.arange(n), np.arange(n)] = 1.0
    return v, np.triu(t), np.triu(a)[:n]


def study():
    rs = np.random.RandomState(0)
    print("panel m x n, condition      | max|V - V_lapack|  max|T - T_lapack|  max|R - R_lapack| / max|R|")
    for m, n, cond in ((4096, 32, 1), (65536, 32, 1), (8192, 32, 1e6), (8192, 32, 1e12), (8192, 64, 1e3)):
        A = rs.randn(m, n)
        if cond > 1:
            u, _, vt = np.linalg.svd(A, full_matrices=False)
            A = (u * np.logspace(0, -np.log10(cond), n)) @ vt
        Q, R = tsqr_explicit_q(A, chunk=max(n, m // 148 + 1))
        Y, T, Rh = householder_from_q(Q, R)
        v, t, r = lapack(A)
        print(f"{m:6d} x {n:3d}, cond {cond:7.0e}   | {np.abs(Y - v).max():.1e}            {np.abs(T - t).max():.1e}            "
              f"{np.abs(Rh - r).max() / np.abs(r).max():.1e}")


if __name__ == "__main__":
    study()
