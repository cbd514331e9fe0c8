"""BigMatrix: a block-sharded N-d array whose tiles live in GPU HBM.

Same constructor, block geometry and get_block / put_block / numpy / free / submatrix / .T
semantics as reference numpywren/matrix.py (BigMatrix :37-560, BigMatrixView :562-782), with the
S3 bucket replaced by an in-process *HBM object store*:

  reference                                   here
  ---------                                   ----
  S3 object  "{prefix}{key}/{s}_{e}_{shard}_" one torch CUDA tensor per tile, keyed by block index
  JSON header object                          header dict in the store (same fields)
  HEAD / GET / np.load                        dict lookup; the tensor itself (no serialisation)
  PUT / np.save                               device-to-device copy (or H2D for host arrays)
  any worker can read any key                 any stream of the owning process; peers through the
                                              placement layer (parallel.py) over NVLink

Tiles are row-major contiguous tensors of the block's shape, so a 4096x4096 fp64 tile is one
128 MiB allocation that TMA can address directly.
"""
from __future__ import annotations

import asyncio
import inspect
import itertools
import os
from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch

from . import utils

DEFAULT_BUCKET = os.environ.get("NPW_B200_BUCKET", "hbm")
DEFAULT_REGION = "local"

_TORCH_DTYPES = {
    np.dtype("float64"): torch.float64, np.dtype("float32"): torch.float32, np.dtype("float16"): torch.float16,
    np.dtype("int64"): torch.int64, np.dtype("int32"): torch.int32, np.dtype("int16"): torch.int16,
    np.dtype("int8"): torch.int8, np.dtype("uint8"): torch.uint8, np.dtype("bool"): torch.bool,
}


def default_device() -> torch.device:
    """Device new tiles are placed on: $NPW_B200_DEVICE, else the current CUDA device.

    There is deliberately no silent CPU default: on a machine without a GPU the caller must ask for
    ``device="cpu"`` (storage-only use, e.g. host-logic tests); the tile kernels refuse CPU tensors.
    """
    env = os.environ.get("NPW_B200_DEVICE")
    if env:
        return torch.device(env)
    if torch.cuda.is_available():
        return torch.device("cuda", torch.cuda.current_device())
    raise RuntimeError("numpywren_b200: no CUDA device visible and NPW_B200_DEVICE is not set; "
                       "pass device='cpu' explicitly for storage-only use")


class _HbmObjectStore:
    """Process-wide replacement for the S3 bucket: bucket → key_base → {header, blocks}."""

    def __init__(self):
        self.buckets: Dict[str, Dict[str, Dict[str, Any]]] = {}

    def entry(self, bucket, key_base):
        # "ready": CUDA events of tiles still being uploaded asynchronously; "mirror": write-through host copies
        return self.buckets.setdefault(bucket, {}).setdefault(
            key_base, {"header": None, "blocks": {}, "ready": {}, "mirror": None, "mirror_events": {}})

    def drop(self, bucket, key_base):
        self.buckets.get(bucket, {}).pop(key_base, None)

    def nbytes(self):
        tot = 0
        for b in self.buckets.values():
            for e in b.values():
                for t in e["blocks"].values():
                    tot += t.numel() * t.element_size()
        return tot


STORE = _HbmObjectStore()

_COPY_STREAMS: Dict[int, Tuple["torch.cuda.Stream", "torch.cuda.Stream"]] = {}


def copy_streams(device: torch.device):
    """(upload, download) side streams of a device: host<->HBM tile traffic overlaps the compute streams."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _COPY_STREAMS:
        _COPY_STREAMS[idx] = (torch.cuda.Stream(device=device), torch.cuda.Stream(device=device))
    return _COPY_STREAMS[idx]


def _run_coro(coro):
    loop = asyncio.new_event_loop()
    try:
        return loop.run_until_complete(coro)
    finally:
        loop.close()


class BigMatrix(object):
    """A multidimensional array stored in HBM, sharded in blocks of a given size.

    Parameters follow reference matrix.py:42-74.  ``bucket``/``prefix``/``region`` are kept for
    signature compatibility: ``bucket`` and ``prefix + key`` name the entry in the HBM object store,
    so two BigMatrix objects built with the same key see the same tiles, like two handles on one
    S3 prefix.  Extra keyword ``device``: where tiles are placed (default: current CUDA device).
    """

    def __init__(self, key, shape=None, shard_sizes=None, bucket=DEFAULT_BUCKET, prefix='numpywren.objects/',
                 dtype=np.float64, parent_fn=None, write_header=False, autosqueeze=True, lambdav=0.0,
                 region=DEFAULT_REGION, safe=True, device=None):
        if bucket is None:
            bucket = os.environ.get('PYWREN_LINALG_BUCKET')
            if bucket is None:
                raise Exception("Bucket not provided and environment variable PYWREN_LINALG_BUCKET not provided.")
        self.bucket = bucket
        self.safe = safe
        self.prefix = prefix
        self.key = key
        self.key_base = os.path.join(prefix, self.key)
        self.dtype = dtype
        self.parent_fn = parent_fn
        self.transposed = False
        self.autosqueeze = autosqueeze
        self.lambdav = lambdav
        self.region = region
        self._device = torch.device(device) if device is not None else None
        header = None
        if shape is None or shard_sizes is None:
            header = self.__read_header__()
        if header is None and shape is None:
            raise Exception("Header doesn't exist and no shape provided.")
        elif shape is None:
            self.shard_sizes = tuple(header['shard_sizes'])
            self.shape = tuple(header['shape'])
            self.dtype = header['dtype']
        else:
            self.shape = tuple(int(x) for x in shape)
            self.shard_sizes = None if shard_sizes is None else tuple(int(x) for x in shard_sizes)
        if (self.shard_sizes is None) or (len(self.shape) != len(self.shard_sizes)):
            raise Exception("shard_sizes should be same length as shape.")
        self.symmetric = False
        if write_header:
            self.__write_header__()
        if (self.lambdav != 0 and (len(self.shape) < 2 or len(set(self.shape)) != 1)):
            raise Exception("Lambda can only be prescribed for square matrices/tensors")

    # ------------------------------------------------------------------ placement
    @property
    def device(self) -> torch.device:
        if self._device is None:
            self._device = default_device()
        return self._device

    @property
    def torch_dtype(self):
        return _TORCH_DTYPES[np.dtype(self.dtype)]

    @property
    def _blocks_store(self) -> Dict[Tuple[int, ...], torch.Tensor]:
        return STORE.entry(self.bucket, self.key_base)["blocks"]

    # ------------------------------------------------------------------ views
    def submatrix(self, *block_slices):
        """Block-sliced view on the same storage (reference matrix.py:133-154)."""
        return BigMatrixView(self, [utils.convert_to_slice(s) for s in block_slices])

    @property
    def T(self):
        """Transposed view on the same storage (reference matrix.py:159-162)."""
        return BigMatrixView(self, [slice(None, None, None)] * len(self.shape), transposed=True)

    # ------------------------------------------------------------------ geometry (matrix.py:426-455, 481-489)
    def _blocks(self, axis=None):
        all_blocks = []
        for i in range(len(self.shape)):
            blocks_axis = [(j, j + self.shard_sizes[i]) for j in range(0, self.shape[i], self.shard_sizes[i])]
            if blocks_axis and blocks_axis[-1][1] > self.shape[i]:
                blocks_axis.pop()
            last_end = blocks_axis[-1][1] if blocks_axis else 0
            if last_end < self.shape[i]:
                blocks_axis.append((last_end, self.shape[i]))
            all_blocks.append(blocks_axis)
        if axis is None:
            return list(itertools.product(*all_blocks))
        elif type(axis) is not int:
            raise Exception("Axis must be an integer.")
        return all_blocks[axis]

    def _block_idxs(self, axis=None):
        idxs = [list(range(len(self._blocks(axis=i)))) for i in range(len(self.shape))]
        if axis is None:
            return list(itertools.product(*idxs))
        elif type(axis) is not int:
            raise Exception("Axis must be integer")
        return idxs[axis]

    def num_blocks(self, axis=None):
        return len(self._block_idxs(axis=axis))

    @property
    def blocks(self):
        return self._blocks()

    @property
    def block_idxs(self):
        return self._block_idxs()

    @property
    def block_idxs_exist(self):
        have = self._blocks_store
        return [b for b in self.block_idxs if b in have]

    @property
    def block_idxs_not_exist(self):
        have = self._blocks_store
        return [b for b in self.block_idxs if b not in have]

    @property
    def blocks_exist(self):
        return [self.__block_idx_to_real_idx__(b) for b in self.block_idxs_exist]

    @property
    def blocks_not_exist(self):
        return [self.__block_idx_to_real_idx__(b) for b in self.block_idxs_not_exist]

    def true_block_idx(self, *block_idx):
        return block_idx

    def __block_idx_to_real_idx__(self, block_idx):
        starts, ends = [], []
        for i in range(len(self.shape)):
            start = block_idx[i] * self.shard_sizes[i]
            end = min(start + self.shard_sizes[i], self.shape[i])
            starts.append(start)
            ends.append(end)
        return tuple(zip(starts, ends))

    def block_shape(self, *block_idx):
        return tuple(e - s for s, e in self.__block_idx_to_real_idx__(block_idx))

    def __shard_idx_to_key__(self, block_idx):
        """The object name the reference would use for this tile (matrix.py:457-464, 491-495)."""
        key_string = ""
        for ((sidx, eidx), shard_size) in zip(self.__block_idx_to_real_idx__(block_idx), self.shard_sizes):
            key_string += "{0}_{1}_{2}_".format(sidx, eidx, shard_size)
        return os.path.join(self.key_base, key_string)

    # ------------------------------------------------------------------ tile access
    def _to_tile(self, block) -> torch.Tensor:
        """Anything array-like → a fresh tensor on this matrix's device."""
        if isinstance(block, torch.Tensor):
            return block.to(device=self.device, copy=True).contiguous()
        arr = np.ascontiguousarray(block)
        return torch.from_numpy(arr).to(self.device)

    def _default_block(self, block_idx):
        pf = self.parent_fn
        if pf is None:
            return None
        if inspect.iscoroutinefunction(pf):
            loop = asyncio.new_event_loop()
            try:
                val = loop.run_until_complete(pf(self, loop, *block_idx))
            finally:
                loop.close()
        else:
            val = pf(self, *block_idx)
        return self._to_tile(val) if not (isinstance(val, torch.Tensor) and val.device == self.device) else val

    def _is_local(self, block_idx) -> bool:
        """With several processes (one per GPU) a tile lives only on its owner rank (parallel.ProcessGrid)."""
        from . import parallel
        grid = parallel.current_grid()
        return grid is None or grid.world == 1 or grid.owner(self, block_idx) == grid.rank

    def _get_block_ref(self, *block_idx):
        """Stored tensor itself (no copy, no squeeze, no lambdav) or None.  Scheduler-internal."""
        return self._blocks_store.get(tuple(int(i) for i in block_idx))

    def _put_block_ref(self, tile: torch.Tensor, *block_idx):
        """Adopt ``tile`` as the stored tensor (no copy).  Scheduler-internal."""
        self._blocks_store[tuple(int(i) for i in block_idx)] = tile

    @property
    def _entry(self):
        return STORE.entry(self.bucket, self.key_base)

    def _ready_event(self, *block_idx):
        """Event of an asynchronous upload still in flight for this tile (None once nobody registered one)."""
        return self._entry["ready"].get(tuple(int(i) for i in block_idx))

    def _upload_async(self, host_tile: torch.Tensor, block_idx):
        """Pinned host tile -> new HBM tile on the upload stream; consumers wait on the recorded event."""
        up, _ = copy_streams(self.device)
        with torch.cuda.stream(up):
            tile = torch.empty(host_tile.shape, dtype=host_tile.dtype, device=self.device)
            tile.copy_(host_tile, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(up)
        self._blocks_store[block_idx] = tile
        self._entry["ready"][block_idx] = ev

    def wait_uploads(self):
        """Block until every ``put_block(..., non_blocking=True)`` upload of this matrix has landed in HBM."""
        ready = self._entry["ready"]
        for ev in list(ready.values()):
            ev.synchronize()
        ready.clear()

    # ---- write-through host mirror (the analogue of "the PUT made the tile durable"): every tile stored into this
    # matrix is also copied to pinned host memory on the download stream, overlapping the rest of the program.
    def mirror_to_host(self, buffers=None):
        """Enable the mirror.  ``buffers``: optional {block_idx: pinned CPU tensor}; missing ones are allocated."""
        e = self._entry
        e["mirror"] = dict(buffers) if buffers else {}
        e["mirror_events"] = {}
        return self

    def _after_put(self, block_idx, tile, event=None):
        e = self._entry
        if e["mirror"] is None or tile is None or not tile.is_cuda:
            return
        block_idx = tuple(int(i) for i in block_idx)
        host = e["mirror"].get(block_idx)
        if host is None:
            host = torch.empty(tile.shape, dtype=tile.dtype, pin_memory=True)
            e["mirror"][block_idx] = host
        _, down = copy_streams(tile.device)
        if event is not None:
            down.wait_event(event)
        else:
            down.wait_stream(torch.cuda.current_stream(tile.device))
        with torch.cuda.stream(down):
            host.view(tile.shape).copy_(tile, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(down)
        tile.record_stream(down)
        e["mirror_events"][block_idx] = ev

    def wait_mirror(self):
        """Block until every mirrored tile has landed in host memory; returns {block_idx: pinned tensor}."""
        e = self._entry
        for ev in e["mirror_events"].values():
            ev.synchronize()
        e["mirror_events"] = {}
        return e["mirror"] or {}

    def get_block(self, *block_idx):
        """Tile at ``block_idx`` as a torch tensor on the owning device (reference matrix.py:266-310).

        Missing tile → ``parent_fn`` default or an exception; ``autosqueeze`` drops unit dims; on the
        diagonal tiles of a square matrix ``lambdav`` is added to the tile's diagonal.  The result
        never aliases stored data.
        """
        if len(block_idx) != len(self.shape):
            raise Exception("Get block query does not match shape {0} vs {1}".format(block_idx, self.shape))
        block_idx = tuple(int(i) for i in block_idx)
        stored = self._blocks_store.get(block_idx)
        if stored is not None and stored.is_cuda:
            ev = self._entry["ready"].get(block_idx)
            if ev is not None:
                torch.cuda.current_stream(stored.device).wait_event(ev)
        if stored is None and not self._is_local(block_idx):
            from . import parallel
            raise Exception("tile {0}{1} is owned by rank {2}; use numpy() (collective) or run it through a program".format(
                self.key, list(block_idx), parallel.current_grid().owner(self, block_idx)))
        if stored is None:
            X_block = self._default_block(block_idx)
            if X_block is None:
                raise Exception("Key does {0} not exist, and no parent function prescripted".format(
                    self.__shard_idx_to_key__(block_idx)))
        else:
            X_block = stored.clone()
        if self.autosqueeze:
            X_block = X_block.squeeze()
        if (len(set(block_idx)) == 1 and len(set(self.shape)) == 1 and len(self.shape) != 1 and self.lambdav != 0):
            self._shift_diagonal(X_block)
        return X_block

    def _shift_diagonal(self, tile):
        if tile.dim() != 2:
            raise Exception("lambdav needs 2-D tiles")
        if tile.is_cuda:
            from . import kernels
            kernels.add_diag(tile, self.lambdav)
        else:  # storage-only host device
            tile.diagonal().add_(self.lambdav)

    async def get_block_async(self, loop, *block_idx):
        return self.get_block(*block_idx)

    def put_block(self, block, *block_idx, non_blocking=False):
        """Store a copy of ``block`` (tensor or ndarray) as tile ``block_idx`` (reference matrix.py:312-361).

        Like the reference's synchronous PUT, the copy is complete when this returns: the caller may reuse ``block``.
        ``non_blocking=True`` (an extension; needs a pinned, contiguous host tensor, else it is ignored) starts the
        host-to-HBM copy on the matrix's upload stream and returns at once — consumers wait on the tile's event, and the
        caller must leave ``block`` untouched until ``wait_uploads()`` (or the program that reads the tile) has finished."""
        block_idx = tuple(int(i) for i in block_idx)
        current_shape = self.block_shape(*block_idx)
        shape = tuple(block.shape)
        if self.autosqueeze:
            if list(shape) == [x for x in current_shape if x != 1]:
                block = block.reshape(current_shape)
                shape = current_shape
        if self.safe and shape != current_shape:
            raise Exception("{2} Incompatible block size: {0} vs {1}".format(shape, current_shape, self))
        if not self._is_local(block_idx):
            return None   # SPMD: every rank issues the same put, only the tile's owner stores it
        if (non_blocking and isinstance(block, torch.Tensor) and not block.is_cuda and block.is_pinned()
                and block.is_contiguous() and self.device.type == "cuda"):
            self._upload_async(block.reshape(current_shape), block_idx)   # overlapped H2D, consumers wait on its event
        else:
            self._entry["ready"].pop(block_idx, None)
            self._blocks_store[block_idx] = self._to_tile(block)
        self._after_put(block_idx, self._blocks_store[block_idx])
        return None

    async def put_block_async(self, block, loop=None, *block_idx, no_overwrite=False):
        if no_overwrite and tuple(int(i) for i in block_idx) in self._blocks_store:
            old = self.get_block(*block_idx)
            new = self._to_tile(block).reshape(old.shape)
            assert torch.allclose(old, new)
        return self.put_block(block, *block_idx)

    def delete_block(self, *block_idx):
        block_idx = tuple(int(i) for i in block_idx)
        self._blocks_store.pop(block_idx, None)
        self._entry["ready"].pop(block_idx, None)

    async def delete_block_async(self, loop=None, *block_idx):
        return self.delete_block(*block_idx)

    def free(self):
        """Delete all allocated blocks while leaving the matrix metadata intact."""
        e = self._entry
        e["blocks"].clear()
        e["ready"].clear()
        e["mirror_events"] = {}
        return 0

    def delete(self):
        """Completely remove the matrix (tiles and header) from the store."""
        STORE.drop(self.bucket, self.key_base)
        return 0

    def numpy(self, workers=None):
        """Gather the whole matrix into a host ndarray (reference matrix.py:410-424)."""
        from . import matrix_utils
        return matrix_utils.get_local_matrix(self, workers)

    # ------------------------------------------------------------------ header (matrix.py:466-474, 535-545)
    def __read_header__(self):
        return STORE.buckets.get(self.bucket, {}).get(self.key_base, {}).get("header")

    def __write_header__(self):
        STORE.entry(self.bucket, self.key_base)["header"] = {
            "shape": tuple(self.shape), "shard_sizes": tuple(self.shard_sizes), "dtype": self.dtype}

    def __delete_header__(self):
        e = STORE.buckets.get(self.bucket, {}).get(self.key_base)
        if e is not None:
            e["header"] = None

    def _register_parent(self, parent_fn):
        self.parent_fn = parent_fn

    def __str__(self):
        return "{0}({1})".format(self.__class__.__name__, self.key)


class BigMatrixView(BigMatrix):
    """Block-slice / transpose view (reference matrix.py:562-782): indices are remapped to the parent
    and tiles are transposed on the way in and out."""

    def __init__(self, parent, parent_slices, transposed=False):
        self.parent = parent
        self.transposed = transposed
        self.bucket = parent.bucket
        self.prefix = parent.prefix
        self.key = parent.key
        self.key_base = parent.key_base
        self.dtype = parent.dtype
        self.parent_fn = parent.parent_fn
        self.autosqueeze = parent.autosqueeze
        self.lambdav = parent.lambdav
        self.safe = parent.safe
        self.region = parent.region
        self._device = parent._device
        self.shard_sizes = tuple(parent.shard_sizes)
        self.parent_slices = []
        self.shape = []
        if isinstance(parent_slices, (int, slice)):
            parent_slices = [parent_slices]
        self.axis_lens = [int(np.ceil(parent.shape[i] / self.shard_sizes[i])) for i in range(len(parent.shape))]
        for i, parent_slice in enumerate(parent_slices):
            start = 0 if parent_slice.start is None else parent_slice.start
            stop = self.axis_lens[i] if parent_slice.stop is None else parent_slice.stop
            step = 1 if parent_slice.step is None else parent_slice.step
            self.shape.append(self.shard_sizes[i] * int(np.ceil((stop - start) / step)))
            # the view's last block may be the parent's ragged last block
            if (stop == self.axis_lens[i] and (stop - 1 - start) % step == 0 and
                    parent.shape[i] % self.shard_sizes[i] != 0):
                self.shape[-1] += parent.shape[i] % self.shard_sizes[i] - self.shard_sizes[i]
            self.parent_slices.append(slice(start, stop, step))
        for i in range(len(self.parent_slices), len(parent.shape)):
            self.parent_slices.append(slice(0, self.axis_lens[i], 1))
            self.shape.append(parent.shape[i])
        if self.transposed:
            self.shape = tuple(reversed(self.shape))
            self.shard_sizes = tuple(reversed(self.shard_sizes))
        self.shape = tuple(self.shape)
        assert len(self.shard_sizes) == len(self.shape)

    @property
    def device(self):
        return self.parent.device

    @property
    def _blocks_store(self):
        return self.parent._blocks_store

    def true_block_idx(self, *block_idx):
        return self.parent.true_block_idx(*self.__view_to_parent_block_idx__(block_idx))

    def _transpose_tile(self, block):
        if isinstance(block, torch.Tensor):
            if block.dim() < 2:
                return block
            if block.is_cuda and block.dim() == 2 and block.dtype == torch.float64:
                from . import kernels
                return kernels.transpose(block)
            return block.permute(*reversed(range(block.dim()))).contiguous()
        return np.ascontiguousarray(np.transpose(block))

    def get_block(self, *block_idx):
        block = self.parent.get_block(*self.__view_to_parent_block_idx__(block_idx))
        if self.transposed:
            block = self._transpose_tile(block)
        return block

    async def get_block_async(self, loop, *block_idx):
        return self.get_block(*block_idx)

    def _get_block_ref(self, *block_idx):
        if self.transposed:
            return None  # a transposed view has no stored tensor of its own
        return self.parent._get_block_ref(*self.__view_to_parent_block_idx__(block_idx))

    def _put_block_ref(self, tile, *block_idx):
        if self.transposed:
            tile = self._transpose_tile(tile)
        return self.parent._put_block_ref(tile, *self.__view_to_parent_block_idx__(block_idx))

    def put_block(self, block, *block_idx):
        if self.transposed:
            block = self._transpose_tile(block)
        return self.parent.put_block(block, *self.__view_to_parent_block_idx__(block_idx))

    async def put_block_async(self, block, loop=None, *block_idx):
        return self.put_block(block, *block_idx)

    def delete_block(self, *block_idx):
        return self.parent.delete_block(*self.__view_to_parent_block_idx__(block_idx))

    def free(self):
        for b in self.block_idxs:
            self.delete_block(*b)
        return 0

    def _blocks(self, axis=None):
        if axis is not None:
            n = len(self._block_idxs(axis=axis))
            out = []
            for j in range(n):
                s = j * self.shard_sizes[axis]
                out.append((s, min(s + self.shard_sizes[axis], self.shape[axis])))
            return out
        return list(itertools.product(*[self._blocks(axis=i) for i in range(len(self.shape))]))

    def _block_idxs(self, axis=None):
        if axis is None:
            return list(itertools.product(*[self._block_idxs(axis=i) for i in range(len(self.shape))]))
        parent_axis = self.__view_to_parent_axis__(axis)
        parent_idxs = self.parent._block_idxs(axis=parent_axis)
        valid = [x for x in parent_idxs if self.__is_valid_parent_block_idx__(x, axis=parent_axis)]
        return [self.__parent_to_view_block_idx__(x, axis=parent_axis) for x in valid]

    def __block_idx_to_real_idx__(self, block_idx):
        return tuple(self._blocks(axis=i)[b] for i, b in enumerate(block_idx))

    @property
    def block_idxs_exist(self):
        have = self._blocks_store
        return [b for b in self.block_idxs if self.true_block_idx(*b) in have]

    @property
    def block_idxs_not_exist(self):
        have = self._blocks_store
        return [b for b in self.block_idxs if self.true_block_idx(*b) not in have]

    def __view_to_parent_axis__(self, view_axis):
        if self.transposed:
            view_axis = len(self.shape) - view_axis - 1
        return view_axis

    def __view_to_parent_block_idx__(self, view_idx):
        intermediate_idx = [elt for elt in view_idx]
        if len(view_idx) < len(self.shape):
            for i in range(len(self.shape)):
                if self.shape[i] <= self.shard_sizes[i]:
                    intermediate_idx.insert(i, 0)
        if len(intermediate_idx) != len(self.shape):
            raise ValueError("Invalid index length.")
        if self.transposed:
            intermediate_idx = list(reversed(intermediate_idx))
        parent_idx = []
        for parent_slice, elt in zip(self.parent_slices, intermediate_idx):
            parent_elt = elt * parent_slice.step + parent_slice.start
            if parent_elt < 0:
                raise NotImplementedError
            if parent_elt >= parent_slice.stop:
                raise IndexError("Array index out of bounds.")
            parent_idx.append(parent_elt)
        return tuple(parent_idx)

    def __parent_to_view_block_idx__(self, parent_idx, axis=None):
        if axis is not None:
            parent_slices = [self.parent_slices[axis]]
            parent_idx = [parent_idx]
        else:
            parent_slices = self.parent_slices
        view_idx = [(p - s.start) // s.step for p, s in zip(parent_idx, parent_slices)]
        if axis is not None:
            return view_idx[0]
        if self.transposed:
            return tuple(reversed(view_idx))
        return tuple(view_idx)

    def __is_valid_parent_block_idx__(self, parent_idx, axis=None):
        if axis is not None:
            parent_slices = [self.parent_slices[axis]]
            parent_idx = [parent_idx]
        else:
            parent_slices = self.parent_slices
        for elt, sl in zip(parent_idx, parent_slices):
            if elt < 0:
                raise NotImplementedError("Negative indexing not yet supported.")
            if elt < sl.start or elt >= sl.stop or (elt - sl.start) % sl.step != 0:
                return False
        return True

    def __str__(self):
        rep = self.parent.__str__()
        if self.transposed:
            rep += ".T"
        return rep + str(tuple(self.shape))
