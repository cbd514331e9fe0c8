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
