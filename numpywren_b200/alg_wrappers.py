"""Allocate outputs/intermediates, bind a LambdaPACK program, wrap it in a LambdaPackProgram.

Same call forms and return values as reference numpywren/alg_wrappers.py: ``cholesky`` :16-27,
``tsqr`` :30-47, ``gemm`` :49-65, ``qr`` :67-91, ``bdfac`` :94-118 → ``(program, {"outputs": [...], "intermediates": [...],
"compile_time": seconds})``.
"""
import time

import numpy as np

from . import config as npw_config
from . import lambdapack as lp
from .algs import BDFAC, CHOLESKY, GEMM, GEMM_ACC, QR, TSQR
from .compiler import lpcompile_for_execution
from .matrix import BigMatrix
from .matrix_utils import constant_zeros, constant_zeros_ext


def _place_by_row_block(mat, axis):
    """Owner = (block index along ``axis``) mod world, expressed in the process grid's (row, col) coordinates."""
    from . import parallel
    if getattr(mat, "placement", None) is not None:
        return

    def placement(true_idx):
        grid = parallel.current_grid()
        j = int(true_idx[axis])
        if grid is None:
            return j, 0
        r = j % grid.world
        return r // grid.Q, r % grid.Q
    mat.placement = placement


def place_plain_block_cyclic(mat):
    """Owner of tile (a, b) = (a mod P, b mod Q) — the process grid's default map WITHOUT the per-block-column rotation
    that balances lower-triangular tile sets.  For a full GEMM the rotation buys nothing and makes every rank consume
    every block row of A (a rank's C tiles then span all block rows): at N=131072 / tile 8192 on 8 GPUs that is 256
    remote tiles = 131 GB of inbox per rank, against 128 tiles = 64 GB with the plain map."""
    from . import parallel
    if getattr(mat, "placement", None) is not None:
        return

    def placement(true_idx):
        grid = parallel.current_grid()
        a, b = (int(true_idx[-2]), int(true_idx[-1])) if len(true_idx) >= 2 else (int(true_idx[0]), 0)
        if grid is None:
            return a, b
        return a % grid.P, b % grid.Q          # b < Q: ProcessGrid.owner's rotation term (b // Q) vanishes
    mat.placement = placement


def _place_tsqr_tree(mat):
    """Placement of the TSQR trees (tiles indexed (level, j), j a multiple of 2**level): node k = j / 2**level of its level
    lives on rank k mod world.  Leaves (level 0) stay with their row block (rank j mod world); the merges of every level
    are dealt round-robin over ALL ranks instead of piling up on the ranks that hold the even row blocks (with the plain
    row-block map every merge of a 2-GPU run — and every merge above level 3 of an 8-GPU run — lands on rank 0).  The
    price is one extra 2 MiB R tile over NVLink per merge."""
    from . import parallel

    def placement(true_idx):
        grid = parallel.current_grid()
        level, j = int(true_idx[0]), int(true_idx[1])
        k = j >> level
        if grid is None:
            return k, 0
        r = k % grid.world
        return r // grid.Q, r % grid.Q
    mat.placement = placement


def cholesky(X, truncate=0):
    """Tiled Cholesky of the SPD BigMatrix ``X`` → lower factor ``O`` (unwritten upper tiles read as zeros)."""
    b = X.shard_sizes[0]
    nb = X.num_blocks(1)
    S = BigMatrix("Cholesky.Intermediate({0})".format(X.key), shape=(nb + 1, X.shape[0], X.shape[0]),
                  shard_sizes=(1, b, b), bucket=X.bucket, write_header=True, parent_fn=constant_zeros, device=X.device)
    O = BigMatrix("Cholesky({0})".format(X.key), shape=(X.shape[0], X.shape[0]), shard_sizes=(b, b), bucket=X.bucket,
                  write_header=True, parent_fn=constant_zeros, device=X.device)
    t = time.time()
    p0 = lpcompile_for_execution(CHOLESKY, inputs=["I"], outputs=["O"])
    p1 = p0(O, X, S, int(np.ceil(X.shape[0] / b)), truncate)
    c_time = time.time() - t
    program = lp.LambdaPackProgram(p1, config=npw_config.default())
    return program, {"outputs": [O], "intermediates": [S], "compile_time": c_time}


def tsqr(X, truncate=0):
    """Tall-skinny QR of ``X`` (one block column): returns R/V/T trees; final R is R.get_block(levels, 0)."""
    b_fac = 2
    assert (X.shard_sizes[1] == X.shape[1])
    shard_size = X.shard_sizes[0]
    shard_sizes = X.shard_sizes
    num_tree_levels = max(int(np.ceil(np.log2(X.num_blocks(0)) / np.log2(b_fac))), 1)
    R_sharded = BigMatrix("tsqr_R({0})".format(X.key), shape=(num_tree_levels * shard_size, X.shape[0]),
                          shard_sizes=shard_sizes, bucket=X.bucket, write_header=True, safe=False, device=X.device)
    T_sharded = BigMatrix("tsqr_T({0})".format(X.key), shape=(num_tree_levels * shard_size * b_fac, X.shape[0]),
                          shard_sizes=(shard_size * b_fac, shard_size), bucket=X.bucket, write_header=True, safe=False,
                          device=X.device)
    V_sharded = BigMatrix("tsqr_V({0})".format(X.key), shape=(num_tree_levels * shard_size * b_fac, X.shape[0]),
                          shard_sizes=(shard_size * b_fac, shard_size), bucket=X.bucket, write_header=True, safe=False,
                          device=X.device)
    # multi-GPU placement: leaf j lives on rank j mod world (leaves are embarrassingly parallel), the merges of every tree
    # level are dealt round-robin over the ranks (_place_tsqr_tree); only 2 MiB R factors cross GPUs.  Tile shapes are
    # declared because these matrices are allocated with the reference's loose shapes (safe=False, alg_wrappers.py:36-38)
    n = X.shape[1]
    _place_by_row_block(X, axis=0)
    for mat in (R_sharded, T_sharded, V_sharded):
        _place_tsqr_tree(mat)
    R_sharded.tile_shape = lambda idx: (n, n)
    T_sharded.tile_shape = lambda idx: (n, n)
    V_sharded.tile_shape = lambda idx: (X.block_shape(idx[1], 0)[0], n) if idx[0] == 0 else (2 * n, n)
    t = time.time()
    p0 = lpcompile_for_execution(TSQR, inputs=["A"], outputs=["Rs"])
    p1 = p0(X, V_sharded, T_sharded, R_sharded, X.num_blocks(0))
    c_time = time.time() - t
    program = lp.LambdaPackProgram(p1, config=npw_config.default())
    return program, {"outputs": [R_sharded, V_sharded, T_sharded], "intermediates": [], "compile_time": c_time}


def gemm(A, B):
    """Tiled A @ B as M*N*K tile products plus a 4-ary add tree (the DSL GEMM program)."""
    b_fac = 4
    assert (A.shape[1] == B.shape[0])
    assert (A.shard_sizes[1] == B.shard_sizes[0])
    shard_sizes = (A.shard_sizes[0], B.shard_sizes[1])
    num_tree_levels = max(int(np.ceil(np.log2(A.num_blocks(1)) / np.log2(b_fac))), 1)
    Temp = BigMatrix("matmul_test_Temp({0},{1})".format(A.key, B.key),
                     shape=(A.shape[0], B.shape[1], B.shape[0], num_tree_levels),
                     shard_sizes=[A.shard_sizes[0], B.shard_sizes[1], 1, 1], bucket=A.bucket, write_header=True,
                     safe=False, parent_fn=constant_zeros, device=A.device)
    C_sharded = BigMatrix("matmul_test_C({0},{1})".format(A.key, B.key), shape=(A.shape[0], B.shape[1]),
                          shard_sizes=shard_sizes, bucket=A.bucket, write_header=True, device=A.device)
    t = time.time()
    p0 = lpcompile_for_execution(GEMM, inputs=["A", "B"], outputs=["Out"])
    # (M, N, K) as the reference passes them (alg_wrappers.py:61)
    p1 = p0(A, B, A.num_blocks(0), A.num_blocks(1), B.num_blocks(1), Temp, C_sharded)
    c_time = time.time() - t
    program = lp.LambdaPackProgram(p1, config=npw_config.default())
    return program, {"outputs": [C_sharded], "intermediates": [Temp], "compile_time": c_time}


def gemm_kloop(A, B, out_key=None):
    """Tiled A @ B with the legacy binops.gemm schedule (owner of every output tile accumulates over the reduction index
    in place) as a LambdaPACK program — the multi-GPU path of ``binops.gemm``.  Returns (program, meta) like ``gemm``."""
    assert (A.shape[1] == B.shape[0])
    assert (A.shard_sizes[1] == B.shard_sizes[0])
    K = A.num_blocks(1)
    out_key = out_key or "matmul_kloop_C({0},{1})".format(A.key, B.key)
    Acc = BigMatrix("matmul_kloop_Acc({0},{1})".format(A.key, B.key), shape=(K + 1, A.shape[0], B.shape[1]),
                    shard_sizes=(1, A.shard_sizes[0], B.shard_sizes[1]), bucket=A.bucket, write_header=True, safe=False,
                    device=A.device)
    C = BigMatrix(out_key, shape=(A.shape[0], B.shape[1]), shard_sizes=(A.shard_sizes[0], B.shard_sizes[1]),
                  bucket=A.bucket, write_header=True, device=A.device)
    place_plain_block_cyclic(Acc)
    place_plain_block_cyclic(C)
    t = time.time()
    p0 = lpcompile_for_execution(GEMM_ACC, inputs=["A", "B"], outputs=["Out"])
    p1 = p0(A, B, A.num_blocks(0), B.num_blocks(1), K, Acc, C)
    c_time = time.time() - t
    program = lp.LambdaPackProgram(p1, config=npw_config.default())
    return program, {"outputs": [C], "intermediates": [Acc], "compile_time": c_time}


def _loose(key, X, shape, shard_sizes, parent_fn=None, place_axis=None, tile_shape=None):
    """Intermediate of the QR/BDFAC programs: over-allocated index space, shape checks off (safe=False).

    Multi-GPU placement: 1-D cyclic over ALL ranks by the block index along ``place_axis`` (block column for the QR
    sweeps, block row for the LQ sweeps).  A tree node's update kernels write TWO tiles of one block column (row); the
    engine runs a node where its first output lives, so both outputs must share an owner — a 1-D map guarantees it,
    the 2-D block-cyclic map of the Cholesky would not.  ``tile_shape(idx)`` declares the stored tile's shape where the
    loose allocation's block shape says nothing (receivers size their NVLink inbox slots from it)."""
    m = BigMatrix(key, shape=shape, shard_sizes=shard_sizes, bucket=X.bucket, write_header=True, parent_fn=parent_fn,
                  safe=False, device=X.device)
    if place_axis is not None:
        m.placement = lambda true_idx, ax=place_axis: (0, int(true_idx[ax]))
    if tile_shape is not None:
        m.tile_shape = tile_shape
    return m


def qr(A):
    """Tiled Householder QR of the square BigMatrix ``A`` (algs.QR): outputs [Rs, Vs, Ts]; block row i of R is
    Rs[i, i, 0] (diagonal) and Rs[i, k, 0], k > i.  Matrix names and shapes as reference alg_wrappers.py:67-91."""
    b_fac = 2
    N = A.shape[0]
    N_blocks = A.num_blocks(0)
    shard_size = A.shard_sizes[0]
    num_tree_levels = max(int(np.ceil(np.log2(A.num_blocks(0)) / np.log2(b_fac))), 1) + 1
    sq = lambda idx: (shard_size, shard_size)            # every tile of the QR program is shard x shard
    # Vs/Ts/Rs[j, i, level] belong to panel (block column) i; S[j, k, i, level] to trailing block column k
    Vs = _loose("Vs", A, (2 * N, 2 * N, num_tree_levels), (shard_size, shard_size, 1), constant_zeros, 1, sq)
    Ts = _loose("Ts", A, (2 * N, 2 * N, num_tree_levels), (shard_size, shard_size, 1), constant_zeros, 1, sq)
    Rs = _loose("Rs", A, (2 * N, 2 * N, num_tree_levels), (shard_size, shard_size, 1), constant_zeros, 1, sq)
    Ss = _loose("Ss", A, (2 * N, 2 * N, 2 * N, num_tree_levels * shard_size), (shard_size, shard_size, 1, 1), constant_zeros,
                1, sq)
    t = time.time()
    p0 = lpcompile_for_execution(QR, inputs=["I"], outputs=["Rs"])
    p1 = p0(A, Vs, Ts, Rs, Ss, N_blocks, 0)
    c_time = time.time() - t
    program = lp.LambdaPackProgram(p1, config=npw_config.default())
    return program, {"outputs": [Rs, Vs, Ts], "intermediates": [Ss], "compile_time": c_time}


def bdfac(A, truncate=0):
    """Reduction of the square BigMatrix ``A`` to block-bidiagonal form (algs.BDFAC): outputs [L_LQ, R_QR]; the
    diagonal block of stage i is R_QR[i, top level, i], the super-diagonal block L_LQ[i, top level, i + 1]
    (reference tests/test_alg_correctness.py:262-265).  Allocation as reference alg_wrappers.py:94-118."""
    b_fac = 2
    N = A.shape[0]
    N_blocks = A.num_blocks(0)
    shard_size = A.shard_sizes[0]
    num_tree_levels = max(int(np.ceil(np.log2(A.num_blocks(0)) / np.log2(b_fac))), 1) + 1
    b = shard_size
    sq = lambda idx: (b, b)
    tall = lambda idx: (b, b) if idx[1] == 0 else (2 * b, b)     # V of a tree merge stacks two R factors
    wide = lambda idx: (b, b) if idx[1] == 0 else (b, 2 * b)     # ... and its LQ mirror puts two L factors side by side
    # QR sweep of stage i: factors (index [i, level, j]) live with panel i, S_QR[i, level, j, k] with block column k;
    # LQ sweep: factors ([i, level, k]) with stage i, S_LQ[i, level, j, k] with block row j
    V_QR = _loose("V_QR", A, (2 * N, num_tree_levels, 2 * N), (1, 1, shard_size), None, 0, tall)
    T_QR = _loose("T_QR", A, (2 * N, num_tree_levels, 2 * N), (1, 1, shard_size), None, 0, sq)
    R_QR = _loose("R_QR", A, (2 * N, num_tree_levels, 2 * N), (shard_size, 1, shard_size), constant_zeros, 0, sq)
    S_QR = _loose("S_QR", A, (2 * N, num_tree_levels, 2 * N, 2 * N), (1, 1, shard_size, shard_size), constant_zeros, 3, sq)
    V_LQ = _loose("V_LQ", A, (2 * N, num_tree_levels, 2 * N), (1, 1, shard_size), None, 0, wide)
    T_LQ = _loose("T_LQ", A, (2 * N, num_tree_levels, 2 * N), (1, 1, shard_size), None, 0, sq)
    L_LQ = _loose("L_LQ", A, (2 * N, num_tree_levels, 2 * N), (1, 1, shard_size), constant_zeros_ext, 0, sq)
    S_LQ = _loose("S_LQ", A, (2 * N, num_tree_levels, 2 * N, 2 * N), (1, 1, shard_size, shard_size), constant_zeros_ext, 2, sq)
    t = time.time()
    p0 = lpcompile_for_execution(BDFAC, inputs=["I"], outputs=["R_QR", "L_LQ"])
    p1 = p0(A, V_QR, T_QR, S_QR, R_QR, V_LQ, T_LQ, S_LQ, L_LQ, N_blocks, truncate)
    c_time = time.time() - t
    program = lp.LambdaPackProgram(p1, config=npw_config.default())
    return program, {"outputs": [L_LQ, R_QR], "intermediates": [S_LQ, S_QR, T_QR, V_QR, V_LQ, T_LQ], "compile_time": c_time}
