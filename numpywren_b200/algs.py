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


def GEMM_ACC(A: BigMatrix, B: BigMatrix, M: int, N: int, K: int, Acc: BigMatrix, Out: BigMatrix):
    # the legacy binops.gemm schedule (binops.py:19-33: one task per output tile, serial accumulation over the reduction
    # index) written as a LambdaPACK program: Acc[k, i, j] is output tile (i, j) after k partial products (SSA; the
    # engine accumulates in place).  Not in the reference's algs.py — it exists so that binops.gemm runs on the DAG
    # engine, i.e. across GPUs: the owner of C[i, j] computes it, A[i, k] / B[k, j] tiles travel along process rows /
    # columns (SUMMA's communication pattern, derived from the DAG like every other transfer).
    for i in range(0, M):
        for j in range(0, N):
            Acc[1, i, j] = gemm(A[i, 0], B[0, j])
            for k in range(1, K):
                Acc[k + 1, i, j] = gemm_acc(Acc[k, i, j], A[i, k], B[k, j])
            Out[i, j] = identity(Acc[K, i, j])


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


def QR(I: BigMatrix, Vs: BigMatrix, Ts: BigMatrix, Rs: BigMatrix, S: BigMatrix, N: int, truncate: int):
    # tiled Householder QR (algs.py:182-234): per block column i, a TSQR panel (leaf QRs, then a binary tree of
    # triangular merges), a flat application of the leaf reflectors to the trailing tiles, then the tree reflectors
    # applied pairwise down the same tree; S[j, k, i, level] is trailing tile (j, k) before panel i at a tree level
    b_fac = 2
    N_tree_full = ceiling(log(N) / log(2))
    for j in range(0, N):
        Vs[j, 0, N_tree_full], Ts[j, 0, N_tree_full], Rs[j, 0, N_tree_full] = qr_factor(I[j, 0])
    for level in range(0, N_tree_full):
        for j in range(0, N, 2 ** (level + 1)):
            Vs[j, 0, N_tree_full - level - 1], Ts[j, 0, N_tree_full - level - 1], Rs[j, 0, N_tree_full - level - 1] = qr_factor_triangular(Rs[j, 0, N_tree_full - level], Rs[j + 2 ** level, 0, N_tree_full - level])
    for j in range(0, N):
        for k in range(1, N):
            S[j, k, 1, N_tree_full] = qr_leaf(Vs[j, 0, N_tree_full], Ts[j, 0, N_tree_full], I[j, k])
    for k in range(1, N):
        for level in range(0, N_tree_full):
            for j in range(0, N, 2 ** (level + 1)):
                S[j, k, 1, N_tree_full - 1 - level], S[j + 2 ** level, k, 1, 0] = qr_trailing_update(Vs[j, 0, N_tree_full - 1 - level], Ts[j, 0, N_tree_full - 1 - level], S[j, k, 1, N_tree_full - level], S[j + 2 ** level, k, 1, N_tree_full - level])
    for k in range(1, N):
        Rs[0, k, 0] = identity(S[0, k, 1, 0])
    for i in range(1, N):
        N_tree = ceiling(log(N - i) / log(2))
        for j in range(i, N):
            Vs[j, i, N_tree], Ts[j, i, N_tree], Rs[j, i, N_tree] = qr_factor(S[j, i, i, 0])
        for level in range(0, N_tree):
            for j in range(i, N, 2 ** (level + 1)):
                Vs[j, i, N_tree - level - 1], Ts[j, i, N_tree - level - 1], Rs[j, i, N_tree - level - 1] = qr_factor_triangular(Rs[j, i, N_tree - level], Rs[j + 2 ** level, i, N_tree - level])
        for j in range(i, N):
            for k in range(i + 1, N):
                S[j, k, i + 1, N_tree] = qr_leaf(Vs[j, i, N_tree], Ts[j, i, N_tree], S[j, k, i, 0])
        for k in range(i + 1, N):
            for level in range(0, N_tree):
                for j in range(i, N, 2 ** (level + 1)):
                    S[j, k, i + 1, N_tree - 1 - level], S[j + 2 ** level, k, i + 1, 0] = qr_trailing_update(Vs[j, i, N_tree - 1 - level], Ts[j, i, N_tree - 1 - level], S[j, k, i + 1, N_tree - level], S[j + 2 ** level, k, i + 1, N_tree - level])
        for k in range(i + 1, N):
            Rs[i, k, 0] = identity(S[i, k, i + 1, 0])


def BDFAC(I: BigMatrix, V_QR: BigMatrix, T_QR: BigMatrix, S_QR: BigMatrix, R_QR: BigMatrix, V_LQ: BigMatrix, T_LQ: BigMatrix, S_LQ: BigMatrix, L_LQ: BigMatrix, N: int, truncate: int):
    # reduction to block-bidiagonal form (algs.py:38-179): at stage i a TSQR panel of block column i with the
    # reflectors applied to the trailing block row(s), then the mirrored LQ panel of block row i (columns i+1..);
    # index order of V/T/R/L: (stage, tree level, block); of S_QR/S_LQ: (stage, tree level, row block, column block)
    b_fac = 2
    N_tree_QR_full = ceiling(log(N) / log(2))
    for j in range(0, N):
        V_QR[0, 0, j], T_QR[0, 0, j], R_QR[0, 0, j] = qr_factor(I[j, 0])
        for k in range(1, N):
            S_QR[0, 0, j, k] = qr_leaf(V_QR[0, 0, j], T_QR[0, 0, j], I[j, k])
    for level in range(1, N_tree_QR_full + 1):
        for j in range(0, N, 2 ** level):
            V_QR[0, level, j], T_QR[0, level, j], R_QR[0, level, j] = qr_factor(R_QR[0, level - 1, j], R_QR[0, level - 1, j + 2 ** (level - 1)])
            for k in range(1, N):
                S_QR[0, level, j, k], S_QR[0, N_tree_QR_full, j + 2 ** (level - 1), k] = qr_trailing_update(V_QR[0, level, j], T_QR[0, level, j], S_QR[0, level - 1, j, k], S_QR[0, level - 1, j + 2 ** (level - 1), k])
    N_tree_LQ_full = ceiling(log(N - 1) / log(2))
    for k in range(1, N):
        V_LQ[0, 0, k], T_LQ[0, 0, k], L_LQ[0, 0, k] = lq_factor(S_QR[0, N_tree_QR_full, 0, k])
        for j in range(1, N):
            S_LQ[0, 0, j, k] = lq_leaf(V_LQ[0, 0, k], T_LQ[0, 0, k], S_QR[0, N_tree_QR_full, j, k])
    for level in range(1, N_tree_LQ_full + 1):
        for k in range(1, N, 2 ** level):
            V_LQ[0, level, k], T_LQ[0, level, k], L_LQ[0, level, k] = lq_factor(L_LQ[0, level - 1, k], L_LQ[0, level - 1, k + 2 ** (level - 1)])
            for j in range(1, N):
                S_LQ[0, level, j, k], S_LQ[0, N_tree_LQ_full, j, k + 2 ** (level - 1)] = lq_trailing_update(V_LQ[0, level, k], T_LQ[0, level, k], S_LQ[0, level - 1, j, k], S_LQ[0, level - 1, j, k + 2 ** (level - 1)])
    for i in range(1, N - 1 - truncate):
        N_tree_QR = ceiling(log(N - i) / log(2))
        prev_N_tree_LQ = ceiling(log(N - i) / log(2))
        for j in range(i, N):
            V_QR[i, 0, j], T_QR[i, 0, j], R_QR[i, 0, j] = qr_factor(S_LQ[i - 1, prev_N_tree_LQ, j, i])
            for k in range(i + 1, N):
                S_QR[i, 0, j, k] = qr_leaf(V_QR[i, 0, j], T_QR[i, 0, j], S_LQ[i - 1, prev_N_tree_LQ, j, k])
        for level in range(1, N_tree_QR + 1):
            for j in range(i, N, 2 ** level):
                V_QR[i, level, j], T_QR[i, level, j], R_QR[i, level, j] = qr_factor(R_QR[i, level - 1, j], R_QR[i, level - 1, j + 2 ** (level - 1)])
                for k in range(i + 1, N):
                    S_QR[i, level, j, k], S_QR[i, N_tree_QR, j + 2 ** (level - 1), k] = qr_trailing_update(V_QR[i, level, j], T_QR[i, level, j], S_QR[i, level - 1, j, k], S_QR[i, level - 1, j + 2 ** (level - 1), k])
        N_tree_LQ = ceiling(log(N - i - 1) / log(2))
        for k in range(i + 1, N):
            V_LQ[i, 0, k], T_LQ[i, 0, k], L_LQ[i, 0, k] = lq_factor(S_QR[i, N_tree_QR, i, k])
            for j in range(i + 1, N):
                S_LQ[i, 0, j, k] = lq_leaf(V_LQ[i, 0, k], T_LQ[i, 0, k], S_QR[i, N_tree_QR, j, k])
        for level in range(1, N_tree_LQ + 1):
            for k in range(i + 1, N, 2 ** level):
                V_LQ[i, level, k], T_LQ[i, level, k], L_LQ[i, level, k] = lq_factor(L_LQ[i, level - 1, k], L_LQ[i, level - 1, k + 2 ** (level - 1)])
                for j in range(i + 1, N):
                    S_LQ[i, level, j, k], S_LQ[i, N_tree_LQ, j, k + 2 ** (level - 1)] = lq_trailing_update(V_LQ[i, level, k], T_LQ[i, level, k], S_LQ[i, level - 1, j, k], S_LQ[i, level - 1, j, k + 2 ** (level - 1)])
    V_QR[N - 1, 0, N - 1], T_QR[N - 1, 0, N - 1], R_QR[N - 1, 0, N - 1] = qr_factor(S_LQ[N - 2, 0, N - 1, N - 1])
