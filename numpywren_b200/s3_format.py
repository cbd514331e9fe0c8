"""Interop with the reference's on-disk / on-wire tile format.

A numpywren BigMatrix in S3 is a set of objects under ``{prefix}{key}/``:
  * one object per tile named ``{start}_{end}_{shard}_`` repeated per axis (matrix.py:457-464), whose body is the
    ``np.save`` serialisation of the tile (matrix.py:519-533);
  * a ``header`` object: JSON ``{"shape", "shard_sizes", "dtype"}`` with the dtype base64(pickle) encoded
    (matrix.py:535-555).
``export_matrix`` / ``import_matrix`` write / read exactly that layout in a directory (or any mapping-like store), so
tiles can be exchanged with a real numpywren deployment: sync the directory with ``aws s3 sync`` and the reference's
``BigMatrix(key)`` finds header and tiles.  Tiles stream HBM → pinned host → ``np.save`` and back.
"""
from __future__ import annotations

import base64
import io
import json
import os
import pickle

import numpy as np
import torch

from .matrix import BigMatrix


def encode_dtype(dtype) -> str:
    """matrix.py:547-550."""
    return base64.b64encode(pickle.dumps(dtype)).decode("utf-8")


class _DtypeUnpickler(pickle.Unpickler):
    """The header's ``dtype`` field is a pickle (reference matrix.py:547-555) read from a directory synced from an
    external deployment: resolve nothing but NumPy's dtype constructor and scalar type classes, so that a crafted
    header cannot name an arbitrary callable."""

    def find_class(self, module, name):
        if module in ("numpy", "numpy.core.multiarray", "numpy._core.multiarray", "numpy.core.numerictypes",
                      "numpy._core.numerictypes"):
            obj = getattr(np, name, None)
            if name == "dtype" and obj is np.dtype:
                return obj
            if isinstance(obj, type) and issubclass(obj, np.generic):
                return obj
        raise pickle.UnpicklingError("header dtype field names {0}.{1}: only NumPy dtypes are accepted".format(module, name))


def decode_dtype(enc: str):
    """matrix.py:552-555, restricted to NumPy dtypes (np.float64 the class, or an np.dtype instance)."""
    obj = _DtypeUnpickler(io.BytesIO(base64.b64decode(enc))).load()
    if not (isinstance(obj, np.dtype) or (isinstance(obj, type) and issubclass(obj, np.generic))):
        raise pickle.UnpicklingError("header dtype field does not decode to a NumPy dtype")
    return obj


def tile_object_name(bigm, block_idx) -> str:
    """Object key of one tile relative to the bucket root (matrix.py:457-464, 491-495)."""
    return bigm.__shard_idx_to_key__(tuple(block_idx))


def header_bytes(bigm) -> bytes:
    return json.dumps({"shape": list(bigm.shape), "shard_sizes": list(bigm.shard_sizes),
                       "dtype": encode_dtype(bigm.dtype)}).encode("utf-8")


def tile_bytes(tile) -> bytes:
    """np.save payload of a tile (what the reference PUTs)."""
    arr = tile.detach().cpu().numpy() if isinstance(tile, torch.Tensor) else np.asarray(tile)
    bio = io.BytesIO()
    np.save(bio, arr)
    return bio.getvalue()


def export_matrix(bigm, root: str) -> int:
    """Write header + every stored tile of ``bigm`` under ``root`` in the reference layout.  Returns #tiles written."""
    base = os.path.join(root, bigm.key_base)
    os.makedirs(base, exist_ok=True)
    with open(os.path.join(base, "header"), "wb") as f:
        f.write(header_bytes(bigm))
    n = 0
    for bidx in bigm.block_idxs_exist:
        path = os.path.join(root, tile_object_name(bigm, bidx))
        with open(path, "wb") as f:
            # stored tile as the reference would have PUT it: full block shape (autosqueeze is a read-side transform)
            f.write(tile_bytes(bigm._get_block_ref(*bidx) if not bigm.transposed else bigm.get_block(*bidx)))
        n += 1
    return n


def import_matrix(key: str, root: str, prefix: str = "numpywren.objects/", device=None, **kwargs) -> BigMatrix:
    """Build a BigMatrix from a directory in the reference layout (header required) and load the tiles present."""
    base = os.path.join(root, prefix, key)
    with open(os.path.join(base, "header"), "rb") as f:
        header = json.loads(f.read().decode("utf-8"))
    dtype = decode_dtype(header["dtype"])
    bigm = BigMatrix(key, shape=tuple(header["shape"]), shard_sizes=tuple(header["shard_sizes"]), prefix=prefix,
                     dtype=dtype, write_header=True, device=device, **kwargs)
    for bidx in bigm.block_idxs:
        path = os.path.join(root, tile_object_name(bigm, bidx))
        if os.path.exists(path):
            arr = np.load(path)
            safe, bigm.safe = bigm.safe, False        # stored tiles already have the block shape
            try:
                bigm.put_block(arr, *bidx)
            finally:
                bigm.safe = safe
    return bigm
