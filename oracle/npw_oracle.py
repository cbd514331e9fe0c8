"""CPU ORACLE — test infrastructure, NOT product code.

A NumPy/SciPy restatement of the numpywren LambdaPACK hot path: the tile kernels
(reference numpywren/kernels.py), the BigMatrix block semantics they run on
(numpywren/matrix.py) and the tile programs CHOLESKY / GEMM / TSQR (numpywren/algs.py),
executed in program order on the host.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this module, and only as
the checker or the timed CPU baseline.  Nothing under ``numpywren_b200/`` imports it.

Pinning: ``tests/golden/*.npz`` were produced by running the UNMODIFIED reference
(``/root/reference/numpywren``: algs.py DSL → frontend/compiler → kernels.py, with an
in-memory dict behind BigMatrix.get_block_async/put_block_async) through
``oracle/make_golden.py``; ``tests/test_oracle.py`` checks this restatement against them.
The QR kernels are the exception: the reference's ``fast_qr`` needs f2py modules that only
exist in the authors' S3 bucket (kernels.py:22-40,86-89), so ``qr_factor`` here restates
kernels.py:86-105 with SciPy's LAPACK ``dgeqrt`` (nb = n gives the single n x n T that
``dgeqrt3`` returns).  For that kernel the oracle is pinned only against NumPy's QR
(|R| equality, Q-orthogonality, the reference test's own criterion
tests/test_alg_correctness.py:95-102): "parity unpinned" beyond that.  The same holds for
``fast_qr_triangular`` (f2py ``dtpqrt`` module → SciPy's LAPACK ``dtpqrt``).

QR / BDFAC (``run_qr`` / ``run_bdfac``) exist in two semantics.  ``"reference"`` restates the code as written,
including its work-in-progress placeholders (qr_leaf = ``S0 - V.T @ S0``, kernels.py:160-164; the triangular
merge's V collapsing to the identity, kernels.py:120-122) and is pinned bit for bit against a replay of the
unmodified reference programs (tests/golden/qr_*.npz, bdfac_*.npz).  ``"householder"`` is the intended algorithm,
pinned by the criteria of the reference's own tests (tests/test_alg_correctness.py:160-187, 216-275): R equals
np.linalg.qr's up to row signs; the block-bidiagonal factor has the singular values of the input.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg


# --------------------------------------------------------------------------- kernels.py
def add_matrices(*args):
    """kernels.py:16-20."""
    out = np.zeros(args[0].shape)
    for a in args:
        out += a
    return out


def syrk(s, x, y):
    """kernels.py:212-215 (incl. the allclose short-circuit)."""
    if np.allclose(x, 0) or np.allclose(y, 0):
        return s
    return s - x.dot(y.T)


def chol(x):
    """kernels.py:225-226."""
    return np.linalg.cholesky(x)


def trsm(x, y, lower=False, right=True):
    """kernels.py:254-257: dtrsm(alpha=1, a=x.T, b=y, lower=0, side=1) = y @ inv(x.T)."""
    if np.allclose(y, 0):
        return np.zeros((x.shape[1], y.shape[0]))
    return scipy.linalg.blas.dtrsm(1.0, x.T, y, lower=lower, side=int(right))


def gemm(A, B, transpose_A=False, transpose_B=False):
    """kernels.py:239-244."""
    if transpose_A:
        A = A.T
    if transpose_B:
        B = B.T
    return A.dot(B)


def mul(x, y):
    """kernels.py:233-234."""
    return x * y


def identity(x):
    """kernels.py:236-237."""
    return x


def _larft_forward_columnwise(v, tau):
    """LAPACK dlarft(direct='F', storev='C'): the upper-triangular T of the block reflector I - V T V^T."""
    k = v.shape[1]
    t = np.zeros((k, k))
    for i in range(k):
        t[i, i] = tau[i]
        if i > 0:
            t[:i, i] = -tau[i] * (t[:i, :i] @ (v[:, :i].T @ v[:, i]))
    return t


def slow_qr(x):
    """kernels.py:67-84 (dgeqrf + the f2py dlarft module, restated above): the path fast_qr takes for wide inputs
    (n > m).  Returns v (m x k, unit lower), t (k x k), r (k x n upper trapezoidal), k = min(m, n)."""
    qr, tau, work, info = scipy.linalg.lapack.dgeqrf(a=x)
    if info != 0:
        raise RuntimeError(f"dgeqrf info={info}")
    r = np.triu(qr)
    k = min(x.shape[0], x.shape[1])
    v = np.tril(qr)[:, :k].copy()
    v[np.diag_indices(k)] = 1
    r = r[:r.shape[1], :]
    return v, _larft_forward_columnwise(v, tau), r


def fast_qr(x):
    """kernels.py:86-105 with scipy's dgeqrt standing in for the f2py dgeqrt3 module.

    Returns (V, T, R): V m x k unit-lower-trapezoidal, T n x n upper, R n x n upper,
    Q = I - V T V^T.
    """
    m, n = x.shape
    k = min(m, n)
    if n > m:
        return slow_qr(x)                                     # kernels.py:94-95
    a, t, info = scipy.linalg.lapack.dgeqrt(n, np.asfortranarray(x))
    if info != 0:
        raise RuntimeError(f"dgeqrt info={info}")
    r = np.triu(a)
    v = np.triu(a.T).T.copy()
    v = v[:, :k]
    v[np.diag_indices(min(v.shape[0], v.shape[1]))] = 1
    r = r[:r.shape[1], :]
    return v, np.triu(t), r


def qr_factor(*blocks):
    """kernels.py:127-130."""
    return fast_qr(np.vstack(blocks))


def _blocked_triu(t, nb):
    """Keep the upper triangle of every nb-wide column block of the nb x n array t (LAPACK's blocked T storage)."""
    out = np.zeros_like(t)
    n = t.shape[1]
    for k0 in range(0, n, nb):
        w = min(nb, n - k0)
        out[:w, k0:k0 + w] = np.triu(t[:w, k0:k0 + w])
    return out


def fast_qr_triangular(x0, x1, semantics="reference"):
    """kernels.py:107-124 with scipy's dtpqrt standing in for the f2py dtpqrt module (m = n, l = m, nb = min(n, 32)).

    "reference": literal — ``v = np.triu(x1.T).T`` keeps the LOWER triangle of dtpqrt's upper-triangular V, so with the
    unit diagonal v is the identity; t is the n x n zero array whose first nb rows hold LAPACK's blocked T.
    "householder": v = dtpqrt's V2 (n x n upper triangular), t = the single n x n compact-WY T (nb = n),
    Q = I - [I; V2] T [I; V2]^T.
    """
    m, n = x0.shape
    if semantics == "householder":
        a, b, t, info = scipy.linalg.lapack.dtpqrt(m, n, np.asfortranarray(x0), np.asfortranarray(x1))
        if info != 0:
            raise RuntimeError(f"dtpqrt info={info}")
        return np.triu(b), np.triu(t), np.triu(a)
    nb = min(n, 32)
    a, b, t, info = scipy.linalg.lapack.dtpqrt(m, nb, np.asfortranarray(x0), np.asfortranarray(x1))
    if info != 0:
        raise RuntimeError(f"dtpqrt info={info}")
    tfull = np.zeros((n, n))
    tfull[:nb, :] = _blocked_triu(t, nb)
    r = np.triu(a)
    v = np.triu(b.T).T.copy()
    v[np.diag_indices(min(v.shape[0], v.shape[1]))] = 1
    return v, tfull, r


def qr_factor_triangular(x0, x1, semantics="reference"):
    """kernels.py:132-134."""
    return fast_qr_triangular(x0, x1, semantics)


def lq_factor(*blocks):
    """kernels.py:145-150."""
    if len(blocks) == 2:
        assert blocks[0].shape[0] == blocks[1].shape[0]
    ins = np.hstack(blocks)
    v, t, r = fast_qr(ins.T)
    return v.T, t.T, r.T


def lq_leaf(V, T, S0):
    """kernels.py:154-157."""
    return S0 - S0 @ V.T @ T.T @ V


def qr_leaf(V, T, S0, semantics="reference"):
    """kernels.py:160-164: the code as written returns S0 - V.T @ S0 (the compact-WY line is commented out);
    "householder" is that commented line, (I - V T V^T)^T S0."""
    if semantics == "householder":
        return S0 - (V @ (T.T @ (V.T @ S0)))
    return S0 - (V.T @ S0)


def qr_trailing_update(V, T, S0, S1, semantics="reference"):
    """kernels.py:181-188."""
    if S1 is None:
        return qr_leaf(V, T, S0, semantics), np.zeros(S0.shape)
    V = V[-S0.shape[0]:]
    W = T.T @ (S0 + V.T @ S1)
    S01 = S0 - W
    S11 = S1 - V.dot(W)
    return S01, S11


def lq_trailing_update(V, T, S0, S1=None):
    """kernels.py:199-208."""
    if S1 is None:
        return lq_leaf(V, T, S0), np.zeros(S0.shape)
    V = V[:, -S0.shape[0]:]
    W = (S0 + S1 @ V.T) @ T.T
    S01 = S0 - W
    S11 = S1 - W.dot(V)
    assert S0.shape == S01.shape
    assert S1.shape == S11.shape
    return S01, S11


def syrk_flops(s, x, y):
    """kernels.py:217-221."""
    m, n = x.shape
    z = y.shape[1]
    return 2 * m * n * z + m * z


def chol_flops(x):
    """kernels.py:228-229."""
    return (x.shape[0] ** 3) / 3


def trsm_flops(x, y):
    """kernels.py:259-263."""
    return x.shape[0] * x.shape[1] * y.shape[1]


def gemm_flops(A, B):
    """kernels.py:246-249."""
    m, n = A.shape
    return 2 * m * n * B.shape[1]


# --------------------------------------------------------------------------- matrix.py
class OracleBigMatrix:
    """In-memory restatement of BigMatrix block semantics (matrix.py:77-361, 426-489).

    Blocks are NumPy arrays in a dict keyed by block index.  ``parent_fn`` is a plain
    callable ``(bigm, *block_idx) -> ndarray`` (the reference's is async with a loop
    argument, matrix_utils.py:314-317).
    """

    def __init__(self, key, shape, shard_sizes, dtype=np.float64, parent_fn=None, autosqueeze=True, lambdav=0.0,
                 safe=True):
        if len(shape) != len(shard_sizes):
            raise Exception("shard_sizes should be same length as shape.")  # matrix.py:123-124
        self.key = key
        self.shape = tuple(shape)
        self.shard_sizes = tuple(shard_sizes)
        self.dtype = dtype
        self.parent_fn = parent_fn
        self.autosqueeze = autosqueeze
        self.lambdav = lambdav
        self.safe = safe
        self.store = {}
        if self.lambdav != 0 and (len(self.shape) < 2 or len(set(self.shape)) != 1):
            raise Exception("Lambda can only be prescribed for square matrices/tensors")  # matrix.py:129-130

    # matrix.py:426-443
    def _blocks(self, axis=None):
        import itertools
        all_blocks = []
        for i in range(len(self.shape)):
            ax = [(j, j + self.shard_sizes[i]) for j in range(0, self.shape[i], self.shard_sizes[i])]
            if ax[-1][1] > self.shape[i]:
                ax.pop()
            if not ax or ax[-1][1] < self.shape[i]:
                ax.append((ax[-1][1] if ax else 0, self.shape[i]))
            all_blocks.append(ax)
        if axis is None:
            return list(itertools.product(*all_blocks))
        return all_blocks[axis]

    # matrix.py:448-455
    def _block_idxs(self, axis=None):
        import itertools
        idxs = [list(range(len(self._blocks(axis=i)))) for i in range(len(self.shape))]
        if axis is None:
            return list(itertools.product(*idxs))
        return idxs[axis]

    def num_blocks(self, axis=None):
        return len(self._block_idxs(axis=axis))

    @property
    def block_idxs(self):
        return self._block_idxs()

    @property
    def blocks(self):
        return self._blocks()

    # matrix.py:481-489
    def block_idx_to_real_idx(self, block_idx):
        out = []
        for i in range(len(self.shape)):
            start = block_idx[i] * self.shard_sizes[i]
            end = min(start + self.shard_sizes[i], self.shape[i])
            out.append((start, end))
        return tuple(out)

    # matrix.py:273-310
    def get_block(self, *block_idx):
        if len(block_idx) != len(self.shape):
            raise Exception("Get block query does not match shape {0} vs {1}".format(block_idx, self.shape))
        if block_idx in self.store:
            blk = self.store[block_idx].copy()
        elif self.parent_fn is None:
            raise Exception("Key does {0} not exist, and no parent function prescripted".format(block_idx))
        else:
            blk = self.parent_fn(self, *block_idx)
        if self.autosqueeze:
            blk = np.squeeze(blk)
        if len(set(block_idx)) == 1 and len(set(self.shape)) == 1 and len(self.shape) != 1:
            idxs = np.diag_indices(blk.shape[0])
            blk[idxs] += self.lambdav
        return blk

    # matrix.py:318-361
    def put_block(self, block, *block_idx):
        real = self.block_idx_to_real_idx(block_idx)
        current_shape = tuple(e - s for s, e in real)
        if self.autosqueeze:
            if list(block.shape) == [x for x in current_shape if x != 1]:
                block = block.reshape(current_shape)
        if self.safe and block.shape != current_shape:
            raise Exception("Incompatible block size: {0} vs {1}".format(block.shape, current_shape))
        self.store[block_idx] = np.array(block, copy=True)

    # matrix.py:410-424 → matrix_utils.get_local_matrix
    def numpy(self):
        out = np.zeros(self.shape, dtype=self.dtype)
        for bidx, blk in zip(self._block_idxs(), self._blocks()):
            sl = tuple(slice(s, e) for s, e in blk)
            out[sl] = self.get_block(*bidx).reshape(out[sl].shape)
        return out


def constant_zeros(bigm, *block_idx):
    """matrix_utils.py:314-317."""
    real = bigm.block_idx_to_real_idx(block_idx)
    return np.zeros(tuple(e - s for s, e in real))


def constant_zeros_ext(bigm, *block_idx):
    """matrix_utils.py:319-325: always a shard x shard zero tile."""
    return np.zeros((bigm.shard_sizes[-1], bigm.shard_sizes[-1]))


def shard_matrix(bigm, X_local):
    """matrix_init.py:73-96: host ndarray → blocks."""
    for bidx, blk in zip(bigm.block_idxs, bigm.blocks):
        bigm.put_block(X_local[tuple(slice(s, e) for s, e in blk)], *bidx)
    return bigm


# --------------------------------------------------------------------------- algs.py programs
def run_cholesky(I, truncate=0):
    """algs.CHOLESKY (algs.py:236-249) with alg_wrappers.cholesky's allocation (alg_wrappers.py:16-27).

    Returns (O, S): the factor BigMatrix (upper tiles unwritten → zeros via parent_fn) and the
    SSA intermediate.  Executed in program order (any topological order gives identical bits: every
    tile is written once and each task is a pure function of its inputs).
    """
    b = I.shard_sizes[0]
    n = I.shape[0]
    nb = int(np.ceil(n / b))
    S = OracleBigMatrix(f"Cholesky.Intermediate({I.key})", (nb + 1, n, n), (1, b, b), parent_fn=constant_zeros)
    O = OracleBigMatrix(f"Cholesky({I.key})", (n, n), (b, b), parent_fn=constant_zeros)
    N = nb
    O.put_block(chol(I.get_block(0, 0)), 0, 0)
    for j in range(1, N - truncate):
        O.put_block(trsm(O.get_block(0, 0), I.get_block(j, 0)), j, 0)
        for k in range(1, j + 1):
            S.put_block(syrk(I.get_block(j, k), O.get_block(j, 0), O.get_block(k, 0)), 1, j, k)
    for i in range(1, N - truncate):
        O.put_block(chol(S.get_block(i, i, i)), i, i)
        for j in range(i + 1, N - truncate):
            O.put_block(trsm(O.get_block(i, i), S.get_block(i, j, i)), j, i)
            for k in range(i + 1, j + 1):
                S.put_block(syrk(S.get_block(i, j, k), O.get_block(j, i), O.get_block(k, i)), i + 1, j, k)
    return O, S


def run_gemm(A, B):
    """algs.GEMM (algs.py:251-266) with alg_wrappers.gemm's allocation (alg_wrappers.py:49-65).

    NB the wrapper passes (M, N, K) = (A.nb(0), A.nb(1), B.nb(1)) (alg_wrappers.py:61), which is only
    meaningful for square tile grids; restated literally.
    """
    b_fac = 4
    num_tree_levels = max(int(np.ceil(np.log2(A.num_blocks(1)) / np.log2(b_fac))), 1)
    Temp = OracleBigMatrix("Temp", (A.shape[0], B.shape[1], B.shape[0], num_tree_levels),
                           (A.shard_sizes[0], B.shard_sizes[1], 1, 1), safe=False, parent_fn=constant_zeros)
    Out = OracleBigMatrix("Out", (A.shape[0], B.shape[1]), (A.shard_sizes[0], B.shard_sizes[1]))
    M, N, K = A.num_blocks(0), A.num_blocks(1), B.num_blocks(1)
    tree_depth = int(math.ceil(math.log(K) / math.log(4))) if K > 1 else 0
    for i in range(M):
        for j in range(N):
            for k in range(K):
                Temp.put_block(gemm(A.get_block(i, k), B.get_block(k, j)), i, j, k, 0)
    for i in range(M):
        for j in range(N):
            for level in range(tree_depth):
                for k in range(0, K, 4 ** (level + 1)):
                    Temp.put_block(add_matrices(Temp.get_block(i, j, k, level),
                                                Temp.get_block(i, j, k + 4 ** level, level),
                                                Temp.get_block(i, j, k + 2 * 4 ** level, level),
                                                Temp.get_block(i, j, k + 3 * 4 ** level, level)), i, j, k, level + 1)
    for i in range(M):
        for j in range(N):
            Out.put_block(identity(Temp.get_block(i, j, 0, tree_depth)), i, j)
    return Out, Temp


def run_tsqr(A):
    """algs.TSQR (algs.py:30-36) with alg_wrappers.tsqr's allocation (alg_wrappers.py:30-47).

    Returns (Rs, Vs, Ts); the final R is Rs.get_block(num_levels, 0).
    """
    b_fac = 2
    shard_size = A.shard_sizes[0]
    nblocks = A.num_blocks(0)
    num_tree_levels = max(int(np.ceil(np.log2(nblocks) / np.log2(b_fac))), 1)
    Rs = OracleBigMatrix("R", (num_tree_levels * shard_size, A.shape[0]), A.shard_sizes, safe=False)
    Ts = OracleBigMatrix("T", (num_tree_levels * shard_size * b_fac, A.shape[0]), (shard_size * b_fac, shard_size), safe=False)
    Vs = OracleBigMatrix("V", (num_tree_levels * shard_size * b_fac, A.shape[0]), (shard_size * b_fac, shard_size), safe=False)
    N = nblocks
    for j in range(N):
        v, t, r = qr_factor(A.get_block(j, 0))
        Vs.put_block(v, 0, j); Ts.put_block(t, 0, j); Rs.put_block(r, 0, j)
    levels = int(math.ceil(math.log(N) / math.log(2))) if N > 1 else 0
    for level in range(levels):
        for j in range(0, N, 2 ** (level + 1)):
            v, t, r = qr_factor(Rs.get_block(level, j), Rs.get_block(level, j + 2 ** level))
            Vs.put_block(v, level + 1, j); Ts.put_block(t, level + 1, j); Rs.put_block(r, level + 1, j)
    return Rs, Vs, Ts


def _ceil_log2(x):
    """ceiling(log(x)/log(2)) for a positive integer x, exact (the DSL evaluates it with sympy)."""
    return 0 if x <= 1 else (int(x) - 1).bit_length()


def run_qr(A, semantics="reference"):
    """algs.QR (algs.py:182-234) with alg_wrappers.qr's allocation (alg_wrappers.py:67-91).  Returns (Rs, Vs, Ts, S)."""
    N = A.shape[0]
    nb = A.num_blocks(0)
    b = A.shard_sizes[0]
    ntl = max(int(np.ceil(np.log2(nb) / np.log2(2))), 1) + 1
    mk = lambda key, shape, ss: OracleBigMatrix(key, shape, ss, parent_fn=constant_zeros, safe=False)
    Vs = mk("Vs", (2 * N, 2 * N, ntl), (b, b, 1))
    Ts = mk("Ts", (2 * N, 2 * N, ntl), (b, b, 1))
    Rs = mk("Rs", (2 * N, 2 * N, ntl), (b, b, 1))
    S = mk("Ss", (2 * N, 2 * N, 2 * N, ntl * b), (b, b, 1, 1))
    I = A
    N = nb

    def put3(res, idx):
        Vs.put_block(res[0], *idx); Ts.put_block(res[1], *idx); Rs.put_block(res[2], *idx)

    def panel(i, src, N_tree):
        for j in range(i, N):
            put3(qr_factor(src(j, i)), (j, i, N_tree))
        for level in range(0, N_tree):
            for j in range(i, N, 2 ** (level + 1)):
                put3(qr_factor_triangular(Rs.get_block(j, i, N_tree - level), Rs.get_block(j + 2 ** level, i, N_tree - level),
                                          semantics), (j, i, N_tree - level - 1))
        for j in range(i, N):
            for k in range(i + 1, N):
                S.put_block(qr_leaf(Vs.get_block(j, i, N_tree), Ts.get_block(j, i, N_tree), src(j, k), semantics),
                            j, k, i + 1, N_tree)
        for k in range(i + 1, N):
            for level in range(0, N_tree):
                for j in range(i, N, 2 ** (level + 1)):
                    s01, s11 = qr_trailing_update(Vs.get_block(j, i, N_tree - 1 - level), Ts.get_block(j, i, N_tree - 1 - level),
                                                  S.get_block(j, k, i + 1, N_tree - level),
                                                  S.get_block(j + 2 ** level, k, i + 1, N_tree - level), semantics)
                    S.put_block(s01, j, k, i + 1, N_tree - 1 - level)
                    S.put_block(s11, j + 2 ** level, k, i + 1, 0)
        for k in range(i + 1, N):
            Rs.put_block(identity(S.get_block(i, k, i + 1, 0)), i, k, 0)

    panel(0, lambda j, k: I.get_block(j, k), _ceil_log2(N))
    for i in range(1, N):
        panel(i, lambda j, k, i=i: S.get_block(j, k, i, 0), _ceil_log2(N - i))
    return Rs, Vs, Ts, S


def run_bdfac(A, truncate=0, semantics="reference"):
    """algs.BDFAC (algs.py:38-179) with alg_wrappers.bdfac's allocation (alg_wrappers.py:94-118).
    Returns a dict of the eight matrices by their DSL names."""
    NN = A.shape[0]
    N = A.num_blocks(0)
    b = A.shard_sizes[0]
    ntl = max(int(np.ceil(np.log2(N) / np.log2(2))), 1) + 1
    mk = lambda key, shape, ss, pf=None: OracleBigMatrix(key, shape, ss, parent_fn=pf, safe=False)
    V_QR = mk("V_QR", (2 * NN, ntl, 2 * NN), (1, 1, b)); T_QR = mk("T_QR", (2 * NN, ntl, 2 * NN), (1, 1, b))
    R_QR = mk("R_QR", (2 * NN, ntl, 2 * NN), (b, 1, b), constant_zeros)
    S_QR = mk("S_QR", (2 * NN, ntl, 2 * NN, 2 * NN), (1, 1, b, b), constant_zeros)
    V_LQ = mk("V_LQ", (2 * NN, ntl, 2 * NN), (1, 1, b)); T_LQ = mk("T_LQ", (2 * NN, ntl, 2 * NN), (1, 1, b))
    L_LQ = mk("L_LQ", (2 * NN, ntl, 2 * NN), (1, 1, b), constant_zeros_ext)
    S_LQ = mk("S_LQ", (2 * NN, ntl, 2 * NN, 2 * NN), (1, 1, b, b), constant_zeros_ext)
    I = A

    def qr_step(i, src, N_tree):
        for j in range(i, N):
            v, t, r = qr_factor(src(j, i))
            V_QR.put_block(v, i, 0, j); T_QR.put_block(t, i, 0, j); R_QR.put_block(r, i, 0, j)
            for k in range(i + 1, N):
                S_QR.put_block(qr_leaf(V_QR.get_block(i, 0, j), T_QR.get_block(i, 0, j), src(j, k), semantics), i, 0, j, k)
        for level in range(1, N_tree + 1):
            for j in range(i, N, 2 ** level):
                h = 2 ** (level - 1)
                v, t, r = qr_factor(R_QR.get_block(i, level - 1, j), R_QR.get_block(i, level - 1, j + h))
                V_QR.put_block(v, i, level, j); T_QR.put_block(t, i, level, j); R_QR.put_block(r, i, level, j)
                for k in range(i + 1, N):
                    s01, s11 = qr_trailing_update(V_QR.get_block(i, level, j), T_QR.get_block(i, level, j),
                                                  S_QR.get_block(i, level - 1, j, k), S_QR.get_block(i, level - 1, j + h, k),
                                                  semantics)
                    S_QR.put_block(s01, i, level, j, k)
                    S_QR.put_block(s11, i, N_tree, j + h, k)

    def lq_step(i, N_tree_QR, N_tree_LQ):
        for k in range(i + 1, N):
            v, t, l = lq_factor(S_QR.get_block(i, N_tree_QR, i, k))
            V_LQ.put_block(v, i, 0, k); T_LQ.put_block(t, i, 0, k); L_LQ.put_block(l, i, 0, k)
            for j in range(i + 1, N):
                S_LQ.put_block(lq_leaf(V_LQ.get_block(i, 0, k), T_LQ.get_block(i, 0, k), S_QR.get_block(i, N_tree_QR, j, k)),
                               i, 0, j, k)
        for level in range(1, N_tree_LQ + 1):
            for k in range(i + 1, N, 2 ** level):
                h = 2 ** (level - 1)
                v, t, l = lq_factor(L_LQ.get_block(i, level - 1, k), L_LQ.get_block(i, level - 1, k + h))
                V_LQ.put_block(v, i, level, k); T_LQ.put_block(t, i, level, k); L_LQ.put_block(l, i, level, k)
                for j in range(i + 1, N):
                    s01, s11 = lq_trailing_update(V_LQ.get_block(i, level, k), T_LQ.get_block(i, level, k),
                                                  S_LQ.get_block(i, level - 1, j, k), S_LQ.get_block(i, level - 1, j, k + h))
                    S_LQ.put_block(s01, i, level, j, k)
                    S_LQ.put_block(s11, i, N_tree_LQ, j, k + h)

    qr_step(0, lambda j, k: I.get_block(j, k), _ceil_log2(N))
    lq_step(0, _ceil_log2(N), _ceil_log2(N - 1))
    for i in range(1, N - 1 - truncate):
        N_tree_QR = _ceil_log2(N - i)
        prev_N_tree_LQ = _ceil_log2(N - i)
        qr_step(i, lambda j, k, i=i, p=prev_N_tree_LQ: S_LQ.get_block(i - 1, p, j, k), N_tree_QR)
        lq_step(i, N_tree_QR, _ceil_log2(N - i - 1))
    # the last statement is a DAG node like any other: it runs once its input tile has been written, which a
    # truncated program never does (the reference's node then simply never becomes ready)
    if (N - 2, 0, N - 1, N - 1) in S_LQ.store:
        v, t, r = qr_factor(S_LQ.get_block(N - 2, 0, N - 1, N - 1))
        V_QR.put_block(v, N - 1, 0, N - 1); T_QR.put_block(t, N - 1, 0, N - 1); R_QR.put_block(r, N - 1, 0, N - 1)
    return {"V_QR": V_QR, "T_QR": T_QR, "R_QR": R_QR, "S_QR": S_QR, "V_LQ": V_LQ, "T_LQ": T_LQ, "L_LQ": L_LQ, "S_LQ": S_LQ}


def bdfac_assemble(R_QR, L_LQ, n, b, get=lambda m, *idx: m.get_block(*idx)):
    """The block upper-bidiagonal factor from BDFAC's outputs, as the reference test assembles it
    (tests/test_alg_correctness.py:262-265): diagonal block i = R_QR[i, top QR level of stage i, i], super-diagonal
    block i = L_LQ[i, top LQ level of stage i, i + 1]."""
    N = n // b
    fac = np.zeros((n, n))
    for i in range(N):
        fac[i * b:(i + 1) * b, i * b:(i + 1) * b] = get(R_QR, i, _ceil_log2(N - i), i)
        if i + 1 < N:
            fac[i * b:(i + 1) * b, (i + 1) * b:(i + 2) * b] = get(L_LQ, i, _ceil_log2(N - i - 1), i + 1)
    return fac


def binops_gemm(X, Y):
    """Legacy binops.gemm(local=True) (binops.py:19-33,107-174): owner-computes over output tiles, serial k."""
    XY = OracleBigMatrix("XY", (X.shape[0], Y.shape[1]), (X.shard_sizes[0], Y.shard_sizes[1]))
    for i in X._block_idxs(0):
        for j in Y._block_idxs(1):
            acc = None
            for r in X._block_idxs(1):
                p = X.get_block(i, r).dot(Y.get_block(r, j))
                acc = p if acc is None else acc + p
            XY.put_block(acc, i, j)
    return XY


# --------------------------------------------------------------------------- synthetic inputs
def spd_factor_block(j, b, width=128):
    """Row-block X_j (b x width) of the benchmark's SPD generator: RandomState(j).randn."""
    return np.random.RandomState(j).randn(b, width)


def spd_tile(j, k, b, n, width=128):
    """Tile (j,k) of A = X X^T + n I  (tests/test_failures.py:33-37 recipe, per-row-block seeds)."""
    t = spd_factor_block(j, b, width).dot(spd_factor_block(k, b, width).T)
    if j == k:
        t[np.diag_indices(b)] += n
    return t
