"""Legacy non-DAG matrix multiply: ``binops.gemm(pwex, X, Y, ...)`` (reference numpywren/binops.py:107-174).

The reference maps one pywren task per output tile, each running a serial loop over the reduction index
(``_gemm_remote_0``, binops.py:19-33: ``XY_block += X.get_block(i, r).dot(Y.get_block(r, j))``).  Here the same
owner-computes-over-C-tiles schedule is issued straight onto CUDA streams: one stream per output tile (round robin), the
k-loop accumulates in place through the GEMM core (``C = A.B + C``), tiles are read by reference from HBM.
``pwex`` (the pywren executor) is accepted and ignored: there is no remote fan-out on a single box.

On several GPUs (one process per GPU, ``parallel.init_from_env``) the same schedule runs as the LambdaPACK program
``algs.GEMM_ACC`` on the DAG engine: the owner of C[i, j] accumulates it, and the engine's transfer plan moves A[i, k]
along the process row and B[k, j] along the process column over NVLink — SUMMA's communication, derived from the DAG.
"""
from __future__ import annotations

import hashlib

import numpy as np
import torch

from . import kernels
from .matrix import BigMatrix


def generate_key_name_binop(X, Y, op):
    h = hashlib.sha1("{0}|{1}|{2}".format(X.key, Y.key, op).encode()).hexdigest()
    return "{0}({1})".format(op, h)


def _tile(m, *idx):
    """Tile by reference when the matrix stores it directly (no view transposition / lambdav / default-fill)."""
    ref = m._get_block_ref(*idx)
    shifted = (len(set(idx)) == 1 and len(set(m.shape)) == 1 and len(m.shape) != 1 and m.lambdav != 0)
    if ref is None or shifted:
        return m.get_block(*idx)
    ev = m._ready_event(*m.true_block_idx(*idx))
    if ev is not None:
        torch.cuda.current_stream().wait_event(ev)
    return ref.squeeze() if m.autosqueeze else ref


def gemm(pwex, X, Y, out_bucket=None, tasks_per_job=1, local=False, dtype=np.float64, overwrite=True, gemm_impl=0,
         gemm_chunk_size=16, streams=4):
    """Compute X @ Y into a new BigMatrix and return it (all tiles enqueued; the caller's next read synchronises)."""
    reduce_idxs = Y._block_idxs(axis=0)
    if out_bucket is None:
        out_bucket = X.bucket
    if Y.shard_sizes[0] != X.shard_sizes[1]:
        raise Exception("X dim 1 shard size must match Y dim 0 shard size")
    if gemm_impl != 0:
        raise Exception("GEMM IMPL > 0 only supported for standalone mode pywren")
    root_key = generate_key_name_binop(X, Y, "gemm")
    from . import parallel
    grid = parallel.current_grid()
    if grid is not None and grid.world > 1:
        return _gemm_distributed(X, Y, root_key, streams)
    XY = BigMatrix(root_key, shape=(X.shape[0], Y.shape[1]), bucket=out_bucket,
                   shard_sizes=[X.shard_sizes[0], Y.shard_sizes[1]], dtype=dtype, write_header=True, device=X.device)
    todo = list(XY.block_idxs) if overwrite else list(XY.block_idxs_not_exist)
    dev = X.device
    if dev.type != "cuda":
        raise kernels._capi.NpwError(f"binops.gemm: tiles live on {dev}; there is no CPU execution path")
    entry = torch.cuda.Event()
    entry.record(torch.cuda.current_stream(dev))
    pool = [torch.cuda.Stream(device=dev) for _ in range(max(1, min(streams, len(todo))))]
    for n, (i, j) in enumerate(todo):
        if not XY._is_local((i, j)):
            continue
        s = pool[n % len(pool)]
        s.wait_event(entry)
        with torch.cuda.stream(s):
            acc = None
            for r in reduce_idxs:
                a, b = _tile(X, i, r), _tile(Y, r, j)
                if acc is None:
                    acc = kernels.gemm(a, b)
                else:
                    kernels._gemm_accumulate(acc, a, b)
            XY._put_block_ref(acc, i, j)
            done = torch.cuda.Event()
            done.record(s)
            XY._entry["ready"][(i, j)] = done       # readers on other streams wait for this tile's last update
    return XY


def _gemm_distributed(X, Y, root_key, streams):
    """binops.gemm across GPUs: compile algs.GEMM_ACC for (X, Y), run it to completion on the engine, return C."""
    from . import job_runner
    from . import lambdapack as lp
    from .alg_wrappers import gemm_kloop
    program, meta = gemm_kloop(X, Y, out_key=root_key)
    program.start()
    job_runner.lambdapack_run(program, timeout=3600, streams=max(streams, 8))
    if program.program_status() != lp.PS.SUCCESS:
        raise Exception("binops.gemm: program ended with status {0}".format(program.program_status()))
    for m in meta["intermediates"]:
        m.free()
    return meta["outputs"][0]
