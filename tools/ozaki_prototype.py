"""Numerical prototype (CPU, NumPy) of fp64 GEMM emulation on int8 tensor cores — the Ozaki scheme — for the
LambdaPACK `syrk` tile update C = S - X Y^T.  Design study for a tcgen05 `kind::i8` kernel; nothing
here is product code and nothing in numpywren_b200/ imports it.

Scheme (Ootomo, Ozaki, Yokota: "DGEMM on integer matrix multiplication unit", 2024), as it would map to B200:
  1. split   : every row of X (and of Y) is scaled by 2^-e_i (e_i = exponent of the row's largest entry) and cut into s
               signed int8 digit matrices: X = diag(2^e) * sum_p 2^-w_p X_p, widths 6, 7, 7, ... bits (w_p cumulative);
               HBM-bound, done ONCE per panel tile and reused by every syrk of that tile's row / column.
  2. multiply: P_d = sum_{p+q=d} X_p Y_q^T in int32 — exact: |entries| <= k * 2^12 * (#pairs in the group) < 2^31 for
               k = 4096 and up to 64 pairs.  These are the tcgen05.mma kind::i8 products, accumulated per group d in TMEM.
  3. combine : C = S - diag(2^e) (sum_d 2^-(w-weights of d) P_d) diag(2^f) in fp64 (one FMA per group and element).
Pairs with p + q > s + 1 are dropped (their weight is below the last kept digit).

  python tools/ozaki_prototype.py            # prints the accuracy table recorded in DESIGN.md §8
"""
import sys

import numpy as np


def split_rows(A, s):
    """-> (digits [s, m, k] int8, exponents e [m], weights w [s]) with A ~= 2^e[:,None] * sum_p 2^-w[p] * digits[p]."""
    A = np.asarray(A, dtype=np.float64)
    amax = np.abs(A).max(axis=1)
    # e = ceil(log2(amax)) computed exactly from the binary representation (frexp: amax = m 2^ex, m in [0.5, 1)) — the
    # same rule as split_i8_kernel in csrc/npw_ozaki_i8.cu, so that the two can be compared bit for bit
    m, ex = np.frexp(amax)
    e = np.where(amax > 0, np.where(m == 0.5, ex - 1, ex), 0).astype(np.float64)
    r = A / np.exp2(e)[:, None]
    digits = np.zeros((s,) + A.shape, dtype=np.int8)
    w = np.zeros(s)
    bits = 0
    for p in range(s):
        b = 6 if p == 0 else 7
        bits += b
        r = r * (1 << b)
        q = np.rint(r)
        assert np.abs(q).max() <= 64 + (p == 0) * 0, np.abs(q).max()
        digits[p] = q.astype(np.int8)
        r = r - q
        w[p] = bits
    return digits, e, w


def ozaki_gemm_nt(X, Y, s, full=False):
    """X @ Y.T through s int8 digits per operand; int32-exact group sums; pairs with p + q > s + 1 dropped unless full."""
    dx, ex, wx = split_rows(X, s)
    dy, ey, wy = split_rows(Y, s)
    m, n = X.shape[0], Y.shape[0]
    acc = np.zeros((m, n))
    nprod = 0
    # group by d = p + q (0-based: d = 0 .. 2s-2); weight of pair (p, q) is 2^-(wx[p] + wy[q]) and depends on d only
    # because the digit widths are the same sequence on both sides *except* for the first digit — handle exactly:
    for d in range(2 * s - 1):
        if not full and d > s - 1:
            break
        groups = {}
        for p in range(max(0, d - s + 1), min(s, d + 1)):
            q = d - p
            wt = wx[p] + wy[q]
            P = dx[p].astype(np.int32) @ dy[q].astype(np.int32).T        # exact in int32 (checked below)
            nprod += 1
            groups[wt] = groups.get(wt, 0) + P.astype(np.int64)
        for wt, P in groups.items():
            assert np.abs(P).max() < 2 ** 31
            acc += P.astype(np.float64) * np.exp2(-wt)
    return acc * np.exp2(ex)[:, None] * np.exp2(ey)[None, :], nprod


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def tile_study(k=1024, m=256):
    print(f"-- one tile product X Y^T, m = n = {m}, k = {k}, entries ~ N(0,1) and with 1e6 row-scale spread")
    rs = np.random.RandomState(0)
    for name, scale in (("randn", None), ("rows scaled 1e-6..1e6", np.exp(rs.uniform(-14, 14, size=m)))):
        X, Y = rs.randn(m, k), rs.randn(m, k)
        if scale is not None:
            X = X * scale[:, None]
        ref = np.asarray(np.dot(X.astype(np.longdouble), Y.T.astype(np.longdouble)), dtype=np.float64)
        base = rel(X @ Y.T, ref)
        row = [f"{name:24s} fp64 dot {base:.1e} |"]
        for s in (4, 5, 6, 7, 8):
            C, nprod = ozaki_gemm_nt(X, Y, s)
            # error measured the way GEMM error bounds are stated: against |X||Y|^T
            bound = np.abs(X) @ np.abs(Y).T
            row.append(f" s={s}: {rel(C, ref):.1e} (max/|X||Y| {np.abs(C - ref).max() / bound.max():.1e}, {nprod} products)")
        print("".join(row))


def cholesky_study(n=1536, b=256):
    """Blocked right-looking Cholesky (the CHOLESKY program's arithmetic) with the syrk products emulated."""
    sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
    from oracle import npw_oracle as orc
    nb = n // b
    A = np.block([[orc.spd_tile(j, k, b, n, width=128) for k in range(nb)] for j in range(nb)])
    Lref = np.linalg.cholesky(A)
    print(f"-- Cholesky N = {n}, tile {b} (benchmark SPD generator, kappa ~ 2): ||L - L_ref||_F / ||L_ref||_F")
    for s in (None, 4, 5, 6, 7, 8):
        S = {(j, k): A[j * b:(j + 1) * b, k * b:(k + 1) * b].copy() for j in range(nb) for k in range(j + 1)}
        L = np.zeros((n, n))
        for i in range(nb):
            Lii = np.linalg.cholesky(S[(i, i)])
            L[i * b:(i + 1) * b, i * b:(i + 1) * b] = Lii
            for j in range(i + 1, nb):
                L[j * b:(j + 1) * b, i * b:(i + 1) * b] = orc.trsm(Lii, S[(j, i)])
            for j in range(i + 1, nb):
                for k in range(i + 1, j + 1):
                    X, Y = L[j * b:(j + 1) * b, i * b:(i + 1) * b], L[k * b:(k + 1) * b, i * b:(i + 1) * b]
                    S[(j, k)] = S[(j, k)] - (X @ Y.T if s is None else ozaki_gemm_nt(X, Y, s)[0])
        print(f"   syrk {'native fp64' if s is None else 'int8 digits s=%d' % s:20s}: {rel(L, Lref):.2e}")


if __name__ == "__main__":
    tile_study()
    cholesky_study()
