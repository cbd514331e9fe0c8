"""LambdaPACK tile programs on the hot path (the program surface of reference numpywren/algs.py).

These functions are *parsed*, never executed as Python (frontend.parse): every assignment of the
form ``M[...] = kernel(...)`` is one abstract remote call; loop nests span its instances.  The
loop structure and index maps are the reference's, because they define the tile DAG:
CHOLESKY algs.py:236-249, GEMM algs.py:251-266, TSQR algs.py:30-36, and the three compiler
test programs algs.py:3-28.
"""
from numpywren_b200.matrix import BigMatrix


def CHOLESKY(O: BigMatrix, I: BigMatrix, S: BigMatrix, N: int, truncate: int):
    # right-looking tiled Cholesky in SSA form: S[i, j, k] is trailing tile (j, k) after i panel updates
    O[0, 0] = chol(I[0, 0])
    for j in range(1, N - truncate):
        O[j, 0] = trsm(O[0, 0], I[j, 0])
        for k in range(1, j + 1):
            S[1, j, k] = syrk(I[j, k], O[j, 0], O[k, 0])
    for i in range(1, N - truncate):
        O[i, i] = chol(S[i, i, i])
        for j in range(i + 1, N - truncate):
            O[j, i] = trsm(O[i, i], S[i, j, i])
            for k in range(i + 1, j + 1):
                S[i + 1, j, k] = syrk(S[i, j, k], O[j, i], O[k, i])


def GEMM(A: BigMatrix, B: BigMatrix, M: int, N: int, K: int, Temp: BigMatrix, Out: BigMatrix):
    # M*N*K independent tile products, then a 4-ary add tree over k, then a copy-out
    tree_depth = ceiling(log(K) / log(4))
    for i in range(0, M):
        for j in range(0, N):
            for k in range(0, K):
                Temp[i, j, k, 0] = gemm(A[i, k], B[k, j])
    for i in range(0, M):
        for j in range(0, N):
            for level in range(0, tree_depth):
                for k in range(0, K, 4 ** (level + 1)):
                    Temp[i, j, k, level + 1] = add_matrices(Temp[i, j, k, level], Temp[i, j, k + 4 ** level, level], Temp[i, j, k + 2 * 4 ** level, level], Temp[i, j, k + 3 * 4 ** level, level])
    for i in range(0, M):
        for j in range(0, N):
            Out[i, j] = identity(Temp[i, j, 0, tree_depth])


def TSQR(A: BigMatrix, Vs: BigMatrix, Ts: BigMatrix, Rs: BigMatrix, N: int):
    # leaf QR of every row block, then a binary tree of QRs of stacked R factors
    for j in range(0, N):
        Vs[0, j], Ts[0, j], Rs[0, j] = qr_factor(A[j, 0])
    for level in range(0, ceiling(log(N) / log(2))):
        for j in range(0, N, 2 ** (level + 1)):
            Vs[level + 1, j], Ts[level + 1, j], Rs[level + 1, j] = qr_factor(Rs[level, j], Rs[level, j + 2 ** level])


def SimpleTestLinear(A: BigMatrix, B: BigMatrix, N: int):
    for i in range(N):
        for j in range(i + 1, N):
            A[j, i] = identity(A[i, j])
    for z in range(N):
        for k in range(N):
            B[z, k] = identity(A[z, k])


def SimpleTestLinear2(A: BigMatrix, B: BigMatrix, N: int):
    for i in range(N):
        for j in range(i + 1, N):
            A[j + 1, i + j] = identity(A[i, j])
    for z in range(N):
        for k in range(N):
            B[z, k] = identity(A[z, k])


def SimpleTestNonLinear(A: BigMatrix, B: BigMatrix, N: int):
    for i in range(N):
        N_tree = ceiling(log(N - i) / log(2))
        for level in range(0, ceiling(log(N - i) / log(2))):
            for k in range(0, N, 2 ** (level + 1)):
                A[N_tree - level - 1, i, k] = add_matrices(A[N_tree - level, i, k], A[N_tree - level, i, k + 2 ** level])
        B[i] = identity(A[1, i, 0])
