"""Storage helpers: default-fill parent functions and the whole-matrix gather.

Reference numpywren/matrix_utils.py: ``constant_zeros`` :314-317, ``constant_zeros_ext`` :319-325,
``make_constant_parent`` :306-311, ``get_local_matrix`` :156-167 (an mmap + process pool over S3
GETs there; a stream of device→host copies here).
"""
from __future__ import annotations

import numpy as np
import torch


def _block_shape(bigm, block_idx):
    return tuple(e - s for s, e in bigm.__block_idx_to_real_idx__(block_idx))


async def constant_zeros(bigm, loop, *block_idx):
    """parent_fn: an unwritten tile reads as zeros (allocated directly in HBM)."""
    return torch.zeros(_block_shape(bigm, block_idx), dtype=bigm.torch_dtype, device=bigm.device)


async def constant_zeros_ext(bigm, loop, *block_idx):
    """parent_fn used by BDFAC's L/S matrices: always a square shard x shard zero tile."""
    s = bigm.shard_sizes[-1]
    return torch.zeros((s, s), dtype=bigm.torch_dtype, device=bigm.device)


def make_constant_parent(cnst):
    def constant_parent(bigm, *block_idx):
        return torch.full(_block_shape(bigm, block_idx), cnst, dtype=bigm.torch_dtype, device=bigm.device)
    return constant_parent


def get_local_matrix(bigm, workers=None):
    """All tiles → one host ndarray.  Tiles that were never written come from ``parent_fn``."""
    from . import parallel
    grid = parallel.current_grid()
    if grid is not None and grid.world > 1:
        return parallel.gather_numpy(bigm)
    out = np.zeros(tuple(bigm.shape), dtype=bigm.dtype)
    tout = torch.from_numpy(out)
    for bidx, blk in zip(bigm._block_idxs(), bigm._blocks()):
        sl = tuple(slice(s, e) for s, e in blk)
        tile = bigm.get_block(*bidx)
        tout[sl].copy_(tile.reshape(tout[sl].shape))
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    return out


def get_col(bigm, col):
    """Block-column ``col`` of a 2-D BigMatrix as a host ndarray."""
    return np.vstack([bigm.get_block(r, col).cpu().numpy() for r in bigm._block_idxs(0)])


def get_row(bigm, row):
    """Block-row ``row`` of a 2-D BigMatrix as a host ndarray."""
    return np.hstack([bigm.get_block(row, c).cpu().numpy() for c in bigm._block_idxs(1)])


def chunk(l, n):
    if n == 0:
        return []
    return [l[i:i + n] for i in range(0, len(l), n)]
