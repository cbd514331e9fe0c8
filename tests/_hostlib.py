"""TEST DOUBLE for libnpw_b200's C-ABI on HOST memory (tests only — never imported by numpywren_b200/).

The host side of the QR / LQ kernels (numpywren_b200/qr.py, kernels._gemm_any) is pure sequencing of C-ABI calls:
which operand goes where, with which transpose flag, leading dimension, alpha/beta and aliasing.  That logic can be
checked without a GPU by pointing the same Python code at an object that implements the handful of entry points it
uses (include/npw_b200.h) with NumPy on raw host pointers.  ``install(monkeypatch)`` swaps it in for one test:
CPU tensors are then accepted and every call is recorded in ``HostLib.calls``.  The real library is untouched, and
on a GPU box the `-m gpu` tests exercise the very same Python code against the CUDA kernels.
"""
import ctypes

import numpy as np
import scipy.linalg


def _view(ptr, rows, cols, ld):
    if rows == 0 or cols == 0:
        return np.zeros((rows, cols))
    n = (rows - 1) * ld + cols
    flat = np.ctypeslib.as_array(ctypes.cast(ctypes.c_void_p(int(ptr)), ctypes.POINTER(ctypes.c_double)), shape=(n,))
    return np.lib.stride_tricks.as_strided(flat, shape=(rows, cols), strides=(ld * 8, 8))


class HostLib:
    def __init__(self):
        self.calls = []

    # -- bookkeeping entry points
    def npw_last_error(self):
        return b""

    def npw_launch_count(self):
        return len(self.calls)

    def npw_geqrt_work_bytes(self, m, n):
        return 8

    def npw_tpqrt_work_bytes(self, n):
        return 8

    def npw_trsm_work_bytes(self, m, n):
        return 8

    def npw_trsm_rlt_f64(self, B_out, ldbo, L, ldl, B, ldb, m, n, invdiag, work, stream):
        """include/npw_b200.h: B_out = B inv(tril(L)).T (the strict upper triangle of L is ignored)."""
        self.calls.append(("trsm_rlt", m, n))
        l = np.tril(np.array(_view(L, n, n, ldl)))
        b = np.array(_view(B, m, n, ldb))
        _view(B_out, m, n, ldbo)[...] = scipy.linalg.solve_triangular(l, b.T, lower=True, check_finite=False).T
        return 0

    # -- kernels (semantics: include/npw_b200.h)
    def npw_gemm_f64(self, C, ldc, C0, ldc0, A, lda, transA, B, ldb, transB, m, n, k, alpha, beta, stream):
        self.calls.append(("gemm", m, n, k, transA, transB))
        a = _view(A, k, m, lda).T if transA else _view(A, m, k, lda)
        b = _view(B, n, k, ldb).T if transB else _view(B, k, n, ldb)
        acc = alpha * (a @ b)
        if C0 and beta != 0.0:
            acc = acc + beta * _view(C0, m, n, ldc0)
        _view(C, m, n, ldc)[...] = acc
        return 0

    def npw_copy2d_f64(self, dst, ldd, src, lds, rows, cols, trans, stream):
        self.calls.append(("copy2d", rows, cols, trans))
        s = _view(src, rows, cols, lds)
        if trans:
            _view(dst, cols, rows, ldd)[...] = s.T
        else:
            _view(dst, rows, cols, ldd)[...] = s
        return 0

    def npw_fill2d_f64(self, A, lda, rows, cols, mode, value, stream):
        self.calls.append(("fill2d", rows, cols, mode))
        a = _view(A, rows, cols, lda)
        i, j = np.indices((rows, cols))
        if mode == 0:
            a[...] = value
        elif mode == 1:
            a[j < i] = value
        else:
            a[j > i] = value
        return 0

    def npw_add_diag_f64(self, A, lda, rows, cols, lambdav, stream):
        self.calls.append(("add_diag", rows, cols))
        a = _view(A, rows, cols, lda)
        d = np.arange(min(rows, cols))
        a[d, d] += lambdav
        return 0

    def npw_addn_f64(self, out, ptrs, count, nelem, stream):
        self.calls.append(("addn", count, nelem))
        acc = np.zeros(nelem)
        for c in range(count):
            acc += _view(ptrs[c], 1, nelem, nelem)[0]
        _view(out, 1, nelem, nelem)[0][...] = acc
        return 0

    def npw_geqrt_f64(self, V, ldv, T, ldt, R, ldr, A, lda, m, n, work, stream):
        self.calls.append(("geqrt", m, n))
        a = np.array(_view(A, m, n, lda))
        qr, t, info = scipy.linalg.lapack.dgeqrt(n, np.asfortranarray(a))
        assert info == 0
        v = np.tril(qr, -1)[:, :n]
        v[np.arange(n), np.arange(n)] = 1.0
        _view(V, m, n, ldv)[...] = v
        _view(T, n, n, ldt)[...] = np.triu(t)
        _view(R, n, n, ldr)[...] = np.triu(qr)[:n]
        return 0


    # -- Cholesky-path entry points (so that the whole stream engine can be exercised on the host, tests/_fakecuda.py)
    def npw_syrk_f64(self, C, ldc, S, lds, X, ldx, Y, ldy, m, n, k, stream):
        self.calls.append(("syrk", m, n, k))
        _view(C, m, n, ldc)[...] = np.array(_view(S, m, n, lds)) - _view(X, m, k, ldx) @ _view(Y, n, k, ldy).T
        return 0

    def npw_syrk_lower_f64(self, C, ldc, S, lds, X, ldx, Y, ldy, m, n, k, stream):
        """Only the 128 x 128 tiles touching the lower triangle are updated; the others keep S (out of place) / are left
        alone (in place) — include/npw_b200.h."""
        self.calls.append(("syrk_lower", m, n, k))
        full = np.array(_view(S, m, n, lds)) - _view(X, m, k, ldx) @ _view(Y, n, k, ldy).T
        i, j = np.indices((m, n))
        low = (j // 128) * 128 <= (i // 128) * 128 + 127
        out = np.array(_view(S, m, n, lds))
        out[low] = full[low]
        _view(C, m, n, ldc)[...] = out
        return 0

    def npw_invdiag_bytes(self, n):
        return 8 * max(1, ((n + 127) // 128) * 128 * 128)

    def npw_potrf_work_bytes(self, n):
        return 8

    def npw_trtri_diag_f64(self, invdiag, L, ldl, n, stream):
        self.calls.append(("trtri_diag", n))
        return 0

    def npw_potrf_l_f64(self, L_out, ldl, A, lda, n, info, invdiag, work, stream):
        self.calls.append(("potrf", n))
        a = np.tril(np.array(_view(A, n, n, lda)))
        a = a + np.tril(a, -1).T
        l, code = scipy.linalg.lapack.dpotrf(a, lower=1)          # info = order of the first non-positive leading minor
        l = np.tril(l)
        if code != 0:
            l = np.eye(n)                   # the CUDA kernel leaves NaNs; any finite stand-in keeps later host calls alive
        _view(L_out, n, n, ldl)[...] = l
        ctypes.cast(ctypes.c_void_p(int(info)), ctypes.POINTER(ctypes.c_int32))[0] = code
        return 0

    def npw_fill_random_f64(self, A, lda, rows, cols, seed, row0, col0, stream):
        """Any reproducible U(-1, 1) fill keyed by (seed, global row, global column) will do for host-logic tests."""
        self.calls.append(("fill_random", rows, cols))
        i, j = np.indices((rows, cols), dtype=np.uint64)
        x = (i + np.uint64(row0)) * np.uint64(0x9E3779B97F4A7C15) + (j + np.uint64(col0)) * np.uint64(0xBF58476D1CE4E5B9) \
            + np.uint64(seed)
        x ^= x >> np.uint64(30); x *= np.uint64(0xBF58476D1CE4E5B9); x ^= x >> np.uint64(27)
        _view(A, rows, cols, lda)[...] = (x >> np.uint64(11)).astype(np.float64) / float(1 << 53) * 2.0 - 1.0
        return 0

    def npw_mul_f64(self, out, x, y, nelem, stream):
        self.calls.append(("mul", nelem))
        _view(out, 1, nelem, nelem)[0][...] = _view(x, 1, nelem, nelem)[0] * _view(y, 1, nelem, nelem)[0]
        return 0

    # -- EXPERIMENTAL int8 emulation entry points, restated from include/npw_b200.h with the NumPy prototype's arithmetic
    def npw_i8_digits_bytes(self, rows, k, ndigits):
        return rows * k * ndigits

    def npw_split_i8_f64(self, digits, exponents, X, ldx, rows, k, ndigits, stream):
        self.calls.append(("split_i8", rows, k, ndigits))
        x = np.array(_view(X, rows, k, ldx))
        amax = np.abs(x).max(axis=1)
        m, ex = np.frexp(amax)
        e = np.where(amax > 0, np.where(m == 0.5, ex - 1, ex), 0).astype(np.int32)
        r = x / np.exp2(e.astype(np.float64))[:, None]
        d8 = np.ctypeslib.as_array(ctypes.cast(ctypes.c_void_p(int(digits)), ctypes.POINTER(ctypes.c_int8)),
                                   shape=(ndigits, rows, k))
        for p in range(ndigits):
            r = r * (64.0 if p == 0 else 128.0)
            q = np.rint(r)
            d8[p] = q.astype(np.int8)
            r = r - q
        np.ctypeslib.as_array(ctypes.cast(ctypes.c_void_p(int(exponents)), ctypes.POINTER(ctypes.c_int32)), shape=(rows,))[...] = e
        return 0

    def npw_syrk_i8emu_f64(self, C, ldc, S, lds, xd, xe, yd, ye, m, n, k, ndigits, lower_only, stream):
        self.calls.append(("syrk_i8emu", m, n, k, ndigits, lower_only))
        if m % 128 or n % 64 or k % 128:
            return -1001
        as8 = lambda ptr, rows: np.ctypeslib.as_array(ctypes.cast(ctypes.c_void_p(int(ptr)), ctypes.POINTER(ctypes.c_int8)),
                                                      shape=(ndigits, rows, k)).astype(np.int64)
        as32 = lambda ptr, rows: np.ctypeslib.as_array(ctypes.cast(ctypes.c_void_p(int(ptr)), ctypes.POINTER(ctypes.c_int32)),
                                                       shape=(rows,)).astype(np.float64)
        X, Y = as8(xd, m), as8(yd, n)
        acc = np.zeros((m, n))
        for d in range(ndigits):
            P = sum(X[p] @ Y[d - p].T for p in range(d + 1))
            assert np.abs(P).max() < 2 ** 31
            acc += P.astype(np.float64) * 2.0 ** -(12 + 7 * d)
        full = np.array(_view(S, m, n, lds)) - acc * np.exp2(as32(xe, m))[:, None] * np.exp2(as32(ye, n))[None, :]
        out = np.array(_view(S, m, n, lds))
        if lower_only:
            i, j = np.indices((m, n))
            low = (j // 64) * 64 <= (i // 128) * 128 + 127
            out[low] = full[low]
        else:
            out = full
        _view(C, m, n, ldc)[...] = out
        return 0

    def npw_tpqrt_f64(self, V2, ldv, T, ldt, R, ldr, R0, ld0, R1, ld1, n, work, stream):
        """include/npw_b200.h: QR of [triu(R0); triu(R1)] → V2 (bottom half of the reflectors), the single n x n T, R —
        restated here with the general LAPACK QR of the explicit stack, the way the CUDA entry point computes it."""
        self.calls.append(("tpqrt", n))
        stack = np.vstack([np.triu(np.array(_view(R0, n, n, ld0))), np.triu(np.array(_view(R1, n, n, ld1)))])
        qr, t, info = scipy.linalg.lapack.dgeqrt(n, np.asfortranarray(stack))
        assert info == 0
        _view(V2, n, n, ldv)[...] = qr[n:]
        _view(T, n, n, ldt)[...] = np.triu(t)
        _view(R, n, n, ldr)[...] = np.triu(qr)[:n]
        return 0


def install(monkeypatch):
    """Route numpywren_b200's kernel wrappers to a HostLib and let them accept CPU tensors, for one test."""
    import torch
    from numpywren_b200 import _capi, kernels, qr
    lib = HostLib()

    def check_tile(t, name):
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name}: expected a torch.Tensor")
        if t.dtype != torch.float64:
            raise TypeError(f"{name}: expected float64, got {t.dtype}")
        if t.dim() != 2:
            raise ValueError(f"{name}: expected a 2-D tile, got shape {tuple(t.shape)}")

    monkeypatch.setattr(_capi, "load", lambda: lib)
    for mod in (kernels, qr):
        monkeypatch.setattr(mod, "_check_tile", check_tile)
        monkeypatch.setattr(mod, "_stream", lambda: 0)
    return lib


def run_in_program_order(program):
    """Execute every node of a compiled LambdaPACK program on host tiles, in source order (a valid schedule), through
    the public BigMatrix.get_block/put_block and the node's bound kernel — the RemoteRead/RemoteCall/RemoteWrite triple
    (reference lambdapack.py:225-384) without the GPU engine.  A node runs iff it is a starter or all of its (>= 1) DAG parents ran — the reference's readiness
    rule (lambdapack.py:568-584); a node none of whose reads is ever written never becomes ready."""
    import torch
    compiled = program.program
    starters = {(int(e), tuple(sorted((str(k), int(v)) for k, v in vv.items()))) for e, vv in compiled.starters}
    done = set()
    ran = 0
    for node in compiled.nodes:
        if node.key not in starters and not (node.parents and all(p in done for p in node.parents)):
            continue
        tiles = [m.get_block(*idx) for m, idx in node.reads]
        args = []
        for kind, j in node.arg_layout:
            if kind == "read":
                args.append(tiles[j])
            elif isinstance(node.scalars[j], float):
                args.append(node.scalars[j])
        res = node.call.compute(*args)
        res = res if isinstance(res, tuple) else (res,)
        assert len(res) == len(node.writes)
        for (m, idx), t in zip(node.writes, res):
            m.put_block(t if isinstance(t, torch.Tensor) else torch.as_tensor(t), *idx)
        done.add(node.nid)
        ran += 1
    return ran
