"""Host ndarray → BigMatrix tiles (reference numpywren/matrix_init.py:22-31, 73-96)."""
from __future__ import annotations

import hashlib

import numpy as np
import torch

from .matrix import BigMatrix


def shard_matrix(bigm, X_local, n_jobs=1, executor=None, overwrite=True):
    """Scatter ``X_local`` (ndarray or torch tensor) into the tiles of ``bigm``.

    The reference memmaps the array and PUTs each tile to S3 from a thread pool; here each tile is one
    host→device copy (pinned staging when the source is a pageable ndarray) onto the tile's owner.
    """
    if overwrite:
        todo = list(zip(bigm.block_idxs, bigm.blocks))
    else:
        todo = list(zip(bigm.block_idxs_not_exist, bigm.blocks_not_exist))
    src = X_local if isinstance(X_local, torch.Tensor) else torch.from_numpy(np.asarray(X_local))
    for bidxs, blocks in todo:
        sl = tuple(slice(s, e) for s, e in blocks)
        bigm.put_block(src[sl], *bidxs)
    return bigm


def local_numpy_init(X_local, shard_sizes, n_jobs=1, symmetric=False, exists=False, executor=None, write_header=False,
                     bucket=None, overwrite=True, device=None):
    key = "local_" + hashlib.sha1(np.ascontiguousarray(X_local).view(np.uint8)).hexdigest()
    kw = {} if bucket is None else {"bucket": bucket}
    bigm = BigMatrix(key, shape=X_local.shape, shard_sizes=shard_sizes, dtype=X_local.dtype, write_header=write_header,
                     device=device, **kw)
    if not exists:
        return shard_matrix(bigm, X_local, overwrite=overwrite)
    return bigm


def empty_result_matrix(X_sharded, function, args, shape=None, shard_sizes=None, symmetric=False, dtype=None, write_header=False):
    """An unwritten BigMatrix keyed by (function, input key, args) — reference matrix_init.py:33-49."""
    dtype = X_sharded.dtype if dtype is None else dtype
    shape = X_sharded.shape if shape is None else shape
    shard_sizes = X_sharded.shard_sizes if shard_sizes is None else shard_sizes
    name = getattr(function, "__name__", repr(function))
    key = hashlib.sha1((name + X_sharded.key + repr(args)).encode()).hexdigest()
    return BigMatrix(key, shape=shape, shard_sizes=shard_sizes, dtype=dtype, write_header=write_header, bucket=X_sharded.bucket,
                     device=X_sharded.device)


def reshard_down(bigm, breakdowns, pwex=None):
    """A new BigMatrix whose shard sizes are ``bigm.shard_sizes / breakdowns``: every tile is cut into
    prod(breakdowns) equal sub-tiles (reference matrix_init.py:100-147).  ``breakdowns = [2, 2]`` replaces a 4 x 4
    tile by four 2 x 2 tiles.  Each source tile is read once; the sub-tiles are strided device-to-device copies on the
    tile's owner, so nothing leaves HBM (``pwex`` — the reference's pywren executor for a parallel reshard — is
    accepted and ignored)."""
    import itertools
    for x, y in zip(bigm.shard_sizes, breakdowns):
        assert x % y == 0
    new_shard_sizes = [int(x / y) for x, y in zip(bigm.shard_sizes, breakdowns)]
    out = BigMatrix("reshard({0},{1})".format(bigm.key, list(breakdowns)), bucket=bigm.bucket, shape=bigm.shape,
                    shard_sizes=new_shard_sizes, dtype=bigm.dtype, device=bigm.device)
    out.autosqueeze = bigm.autosqueeze
    for old_idx in bigm._block_idxs():
        if not bigm._is_local(tuple(old_idx)):
            continue
        real = bigm.__block_idx_to_real_idx__(old_idx)
        squeeze = bigm.autosqueeze
        bigm.autosqueeze = False                      # sub-tile offsets below are in the tile's full-rank shape
        try:
            data = bigm.get_block(*old_idx)
        finally:
            bigm.autosqueeze = squeeze
        ranges = []
        for ax, (s, e) in enumerate(real):
            step = new_shard_sizes[ax]
            ranges.append([(o, min(o + step, e)) for o in range(s, e, step)])
        for sub in itertools.product(*ranges):
            new_idx = tuple(o // new_shard_sizes[ax] for ax, (o, _) in enumerate(sub))
            sl = tuple(slice(o - real[ax][0], e - real[ax][0]) for ax, (o, e) in enumerate(sub))
            out.put_block(data[sl], *new_idx)
    return out
