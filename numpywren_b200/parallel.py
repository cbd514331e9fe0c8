"""Multi-GPU execution: one process per GPU, tiles sharded by block index, panel tiles exchanged over NVLink.

The reference scales by letting any stateless worker read any tile from S3 (README "S3 as distributed
memory"; lambdapack.py:257,319).  On one 8xB200 box the equivalent of the object store is the union of the
GPUs' HBM: every tile has ONE owner rank (2-D block-cyclic over its block index, like ScaLAPACK/SLATE, which
keeps the shrinking trailing matrix of a factorisation balanced), the owner of a node's output tile executes
the node ("owner computes"), and a tile needed by another rank is pushed to it over NVLink/NVSwitch as soon as
it exists.  All ranks walk the SAME expanded DAG in the SAME order (SPMD), so the set and order of transfers is
known on both sides without any control messages: the sender posts an NCCL send right after the producing
kernel, the receiver posts the matching recv, and consumers wait on the receive like on any other tile event.

The exchange step of blocked Cholesky is exactly one pattern: panel tile O[j,i] (and O[i,i]) goes to the owners
of row j / column j of the trailing matrix.  Nothing else crosses GPUs; there is no data-path collective.
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

_GRID: Optional["ProcessGrid"] = None


def factor_grid(world: int) -> Tuple[int, int]:
    """P x Q with P <= Q and P*Q == world, as square as possible (8 -> 2x4, 4 -> 2x2, 2 -> 1x2)."""
    p = int(np.floor(np.sqrt(world)))
    while world % p:
        p -= 1
    return p, world // p


class ProcessGrid:
    """Rank layout and tile ownership."""

    def __init__(self, world: int, rank: int, shape: Optional[Tuple[int, int]] = None):
        self.world, self.rank = int(world), int(rank)
        self.P, self.Q = shape if shape is not None else factor_grid(self.world)
        if self.P * self.Q != self.world:
            raise ValueError(f"grid {self.P}x{self.Q} does not match world size {self.world}")

    def coords(self, matrix, idx) -> Tuple[int, int]:
        """The two block coordinates that decide ownership.  A matrix may carry its own ``placement(idx)``;
        by default: 2-D → (i, j); 3-D (SSA version first, e.g. Cholesky's S[v, j, k]) → (j, k); 4-D (GEMM's
        Temp[i, j, k, level]) → (i, j); 1-D → (i, 0)."""
        true = matrix.true_block_idx(*idx) if hasattr(matrix, "true_block_idx") else tuple(idx)
        fn = getattr(matrix, "placement", None)
        if fn is not None:
            return fn(true)
        nd = len(true)
        if nd == 1:
            return int(true[0]), 0
        if nd == 2:
            return int(true[0]), int(true[1])
        if nd == 3:
            return int(true[1]), int(true[2])
        return int(true[0]), int(true[1])

    def owner(self, matrix, idx) -> int:
        """2-D block-cyclic with the process row rotated by one every Q block columns:
        rank = ((a + b // Q) mod P) * Q + (b mod Q).  The plain (a mod P, b mod Q) map gives the ranks of the last
        process row ~6 % more work on a LOWER-TRIANGULAR tile set (rows j >= k accumulate in the high process rows);
        the rotation spreads that (work imbalance of the 32x32-tile Cholesky: 2x4 grid 5.8 % -> 1.3 %, 2x2 5.2 % -> 0.4 %)
        at the price of a larger broadcast fan-out per panel tile, which NVLink absorbs easily."""
        a, b = self.coords(matrix, idx)
        return ((a + b // self.Q) % self.P) * self.Q + (b % self.Q)

    def is_mine(self, matrix, idx) -> bool:
        return self.owner(matrix, idx) == self.rank

    def owner_of_coords(self, a: int, b: int) -> int:
        """Owner of a 2-D tile (a, b) under the default map (what ``owner`` computes for a plain 2-D BigMatrix)."""
        return ((a + b // self.Q) % self.P) * self.Q + (b % self.Q)


def set_grid(grid: Optional[ProcessGrid]):
    global _GRID
    _GRID = grid


def current_grid() -> Optional[ProcessGrid]:
    return _GRID


def init_from_env(backend: Optional[str] = None) -> ProcessGrid:
    """Join the torchrun rendezvous (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*), bind this process to its GPU and
    install the process grid.  backend defaults to nccl when CUDA is visible, gloo otherwise (CPU host-logic tests)."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        # the engine uses ~6 compute streams plus one send and one receive stream per peer; give every stream its own
        # hardware queue so that a signal wait at the head of one stream can never stall an unrelated one
        os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        # no device_id: with eager communicator init torch serialises unbatched P2P ops on the world
        # communicator; lazily created per-pair communicators let sends to different peers overlap
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    grid = ProcessGrid(world, rank)
    set_grid(grid)
    return grid


# ------------------------------------------------------------------------------------------- transfer plan
class TransferPlan:
    """Who needs which tile, derived from the expanded DAG (identical on every rank).

    For every tile read by a node executed on a rank other than the tile's owner there is exactly one transfer
    (tile, owner → that rank).  ``after_node[nid]`` lists the transfers triggered by node nid's outputs;
    ``before_node[nid]`` lists transfers of pre-existing tiles (program inputs) first needed by node nid.
    """

    def __init__(self, compiled, grid: ProcessGrid):
        from .compiler import _tile_key
        self.grid = grid
        nodes = compiled.nodes
        self.exec_rank = [grid.owner(*n.writes[0]) for n in nodes]
        self.after_node: Dict[int, List[Tuple[Any, Any, Tuple[int, ...], int, int]]] = {}
        self.before_node: Dict[int, List[Tuple[Any, Any, Tuple[int, ...], int, int]]] = {}
        self.last_use: Dict[Tuple[Any, int], int] = {}       # (tile_key, rank) -> last consumer nid on that rank
        seen = set()
        for n in nodes:
            dst = self.exec_rank[n.nid]
            for (m, idx) in n.reads:
                key = _tile_key(m, idx)
                src = grid.owner(m, idx)
                if src == dst:
                    continue
                self.last_use[(key, dst)] = n.nid
                if (key, dst) in seen:
                    continue
                seen.add((key, dst))
                w = compiled.writer_of(m, idx)
                if w is not None:
                    self.after_node.setdefault(w.nid, []).append((key, m, tuple(idx), src, dst))
                else:
                    self.before_node.setdefault(n.nid, []).append((key, m, tuple(idx), src, dst))
        self.num_transfers = len(seen)

    def assign_inbox_slots(self):
        """Fixed inbox slot of every transfer: {(tile_key, dst): slot}, slot size in elements, slots per inbox.
        Computed from the plan alone, hence identical on every rank; slots of one destination never overlap."""
        slot: Dict[Tuple[Any, int], int] = {}
        counts = [0] * self.grid.world
        slot_elems = 1
        for nid in range(len(self.exec_rank)):
            for lst in (self.before_node.get(nid, ()), self.after_node.get(nid, ())):
                for key, m, idx, src, dst in lst:
                    slot[(key, dst)] = counts[dst]
                    counts[dst] += 1
                    slot_elems = max(slot_elems, int(np.prod(_tile_shape(m, idx))))
        return slot, (slot_elems + 15) // 16 * 16, (max(counts) if counts else 0)

    def describe(self, rank: int) -> List[Tuple[str, Any, int]]:
        """Ordered communication script of one rank: [("send"|"recv", tile_key, peer)] — used by the tests to check
        that every send has a matching recv in the same relative order."""
        out = []
        n_nodes = len(self.exec_rank)
        for nid in range(n_nodes):
            for lst in (self.before_node.get(nid, ()),):
                for key, _, _, src, dst in lst:
                    if src == rank:
                        out.append(("send", key, dst))
                    elif dst == rank:
                        out.append(("recv", key, src))
            for key, _, _, src, dst in self.after_node.get(nid, ()):
                if src == rank:
                    out.append(("send", key, dst))
                elif dst == rank:
                    out.append(("recv", key, src))
        return out


def _tile_shape(m, idx):
    """Shape of the tensor stored for tile ``idx``: the block shape, unless the matrix declares otherwise (matrices the
    reference allocates with loose shapes and safe=False, e.g. TSQR's R/T/V)."""
    fn = getattr(m, "tile_shape", None)
    if fn is not None:
        return tuple(int(x) for x in fn(tuple(idx)))
    return tuple(m.block_shape(*idx))


def _wait_upload(m, idx, ref, engine_event, stream):
    """An input tile stored by ``put_block(..., non_blocking=True)`` from pinned host memory may still be arriving on the
    matrix's upload stream: a transfer of such a tile (no engine event yet) must be ordered after that copy, exactly like
    ``TileEngine._read`` does for local consumers."""
    if ref is None or engine_event is not None or not hasattr(m, "_ready_event"):
        return
    up = m._ready_event(*m.true_block_idx(*idx))
    if up is not None:
        stream.wait_event(up)
        ref.record_stream(stream)


class TileExchange:
    """Posts the planned sends/recvs with torch.distributed P2P (NCCL over NVLink) as the engine walks the DAG."""

    def __init__(self, compiled, grid: ProcessGrid):
        self.grid = grid
        self.rank = grid.rank
        self.plan = TransferPlan(compiled, grid)
        self.cache: Dict[Any, Tuple[torch.Tensor, Any]] = {}   # tile_key -> (buffer, recv work)
        self.pending_sends: List[Any] = []
        self.bytes_sent = 0
        self.bytes_received = 0

    def exec_rank(self, node) -> int:
        return self.plan.exec_rank[node.nid]

    def _do(self, transfers, engine, node_order_hint=None):
        import torch.distributed as dist
        for key, m, idx, src, dst in transfers:
            if src == self.rank:
                ref = m._get_block_ref(*idx)
                if ref is None:
                    tile = m.get_block(*idx)       # default (parent_fn) tile of an input matrix
                else:
                    tile = ref
                ev = engine.tile_event.get(key)
                stream = ev[1] if ev is not None else torch.cuda.current_stream()
                _wait_upload(m, idx, ref, ev, stream)
                with torch.cuda.stream(stream):
                    # issued on the producer's stream: NCCL orders the send after the kernel that wrote the tile
                    self.pending_sends.append((dist.isend(tile.contiguous(), dst), tile))
                self.bytes_sent += tile.numel() * tile.element_size()
            elif dst == self.rank:
                shape = _tile_shape(m, idx)
                buf = torch.empty(shape, dtype=m.torch_dtype, device=engine.device or m.device)
                work = dist.irecv(buf, src)
                self.cache[key] = (buf, work)
                self.bytes_received += buf.numel() * buf.element_size()

    def before_node(self, node, engine):
        t = self.plan.before_node.get(node.nid)
        if t:
            self._do(t, engine)

    def after_node(self, node, engine):
        t = self.plan.after_node.get(node.nid)
        if t:
            self._do(t, engine)
        # receive buffers whose last local consumer has been enqueued can go back to the allocator
        for (m, idx) in node.reads:
            from .compiler import _tile_key
            key = _tile_key(m, idx)
            if key in self.cache and self.plan.last_use.get((key, self.rank)) == node.nid:
                del self.cache[key]

    def remote_tile(self, key, stream) -> Optional[torch.Tensor]:
        ent = self.cache.get(key)
        if ent is None:
            return None
        buf, work = ent
        if not work.is_completed():     # a finished receive needs no ordering (and gloo must not be waited on twice)
            with torch.cuda.stream(stream):
                work.wait()             # stream-level wait on the NCCL receive, the host does not block
        buf.record_stream(stream)
        return buf

    def drain(self):
        for work, _ in self.pending_sends:
            work.wait()
        self.pending_sends = []
        self.cache.clear()


_INBOX: Dict[str, Any] = {}


def _symmetric_inbox(nbytes: int, device):
    """A symmetric (peer-mapped) receive buffer shared by all programs of this process: allocated collectively once and
    re-used while it is large enough.  Every rank must call this with the same size."""
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm_mem
    ent = _INBOX.get("inbox")
    if ent is not None and ent[0] >= nbytes:
        return ent[1], ent[2]
    elems = (int(nbytes) + 7) // 8
    buf = symm_mem.empty(elems, dtype=torch.float64, device=device)
    hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
    _INBOX["inbox"] = (elems * 8, buf, hdl)
    return buf, hdl


class SymmTileExchange(TileExchange):
    """Panel tiles are written straight into the consumer's HBM over NVLink (no NCCL, no SMs).

    Every rank owns a symmetric *inbox* (torch symmetric memory: the same allocation is mapped into every peer's address
    space).  The transfer plan gives each (tile, destination) a fixed slot in the destination's inbox, known to both
    sides.  The producer enqueues, on a per-destination side stream: wait(producer event) → cudaMemcpyAsync of the tile
    into the peer-mapped slot (copy engine) → put_signal(dst).  The consumer enqueues on a per-source side stream:
    wait_signal(src) → record event; compute streams wait on that event and then read the tile *in place* from the inbox.
    Signals are per (src, dst) binary semaphores consumed in plan order, so the k-th wait pairs with the k-th copy.
    """

    SIGNAL_TIMEOUT_MS = int(os.environ.get("NPW_B200_SIGNAL_TIMEOUT_MS", "120000"))

    def __init__(self, compiled, grid: ProcessGrid, device):
        super().__init__(compiled, grid)
        self.device = device
        self.slot, self.slot_elems, self.max_slots = self.plan.assign_inbox_slots()
        # inbox slots are typed and sized for fp64 tiles (the only dtype the C-ABI kernels compute in)
        for lst in list(self.plan.before_node.values()) + list(self.plan.after_node.values()):
            for _key, m, _idx, _src, _dst in lst:
                if m.torch_dtype != torch.float64:
                    raise TypeError("SymmTileExchange moves float64 tiles only; {0} is {1} (use NPW_B200_EXCHANGE=nccl)".format(
                        m.key, m.torch_dtype))
        self.inbox, self.hdl = _symmetric_inbox(max(1, self.max_slots) * self.slot_elems * 8, device)
        self.send_streams: Dict[int, torch.cuda.Stream] = {}
        self.recv_streams: Dict[int, torch.cuda.Stream] = {}

    def _stream(self, table, peer):
        s = table.get(peer)
        if s is None:
            s = torch.cuda.Stream(device=self.device, priority=-1)
            table[peer] = s
        return s

    def _do(self, transfers, engine, node_order_hint=None):
        for key, m, idx, src, dst in transfers:
            shape = _tile_shape(m, idx)
            off = self.slot[(key, dst)] * self.slot_elems
            if src == self.rank:
                ref = m._get_block_ref(*idx)
                tile = m.get_block(*idx) if ref is None else ref
                ev = engine.tile_event.get(key)
                s = self._stream(self.send_streams, dst)
                if ev is not None:
                    s.wait_event(ev[0])
                else:
                    s.wait_stream(torch.cuda.current_stream(self.device))
                _wait_upload(m, idx, ref, ev, s)
                with torch.cuda.stream(s):
                    peer = self.hdl.get_buffer(dst, shape, torch.float64, off)
                    peer.copy_(tile.reshape(shape), non_blocking=True)
                    self.hdl.put_signal(dst, 0, self.SIGNAL_TIMEOUT_MS)
                tile.record_stream(s)
                self.bytes_sent += tile.numel() * tile.element_size()
            elif dst == self.rank:
                r = self._stream(self.recv_streams, src)
                with torch.cuda.stream(r):
                    self.hdl.wait_signal(src, 0, self.SIGNAL_TIMEOUT_MS)
                    ev = torch.cuda.Event()
                    ev.record(r)
                n = int(np.prod(shape))
                buf = self.inbox[off:off + n].view(shape)
                self.cache[key] = (buf, ev)
                self.bytes_received += n * 8

    def remote_tile(self, key, stream) -> Optional[torch.Tensor]:
        ent = self.cache.get(key)
        if ent is None:
            return None
        buf, ev = ent
        stream.wait_event(ev)
        return buf

    def drain(self):
        import torch.distributed as dist
        torch.cuda.synchronize(self.device)
        self.cache.clear()
        # nobody may start overwriting inbox slots (next program) before every rank has finished reading them
        dist.barrier(device_ids=[self.device.index])


def make_exchange(compiled, grid: ProcessGrid, device):
    """NVLink tile exchange for this program: peer-memory writes (default) or NCCL send/recv (NPW_B200_EXCHANGE=nccl)."""
    mode = os.environ.get("NPW_B200_EXCHANGE", "symm")
    if mode == "symm":
        return SymmTileExchange(compiled, grid, device)
    return TileExchange(compiled, grid)


# ------------------------------------------------------------------------------------------- collectives on BigMatrix
def gather_numpy(bigm) -> np.ndarray:
    """Collective: every rank receives the whole matrix as a host ndarray (owner broadcasts each tile)."""
    import torch.distributed as dist
    grid = current_grid()
    out = np.zeros(tuple(bigm.shape), dtype=bigm.dtype)
    tout = torch.from_numpy(out)
    dev = bigm.device
    for bidx, blk in zip(bigm._block_idxs(), bigm._blocks()):
        sl = tuple(slice(s, e) for s, e in blk)
        owner = grid.owner(bigm, bidx)
        shape = tuple(e - s for s, e in blk)
        have = torch.zeros(1, dtype=torch.int32, device=dev)
        if owner == grid.rank:
            ref = bigm._get_block_ref(*bidx)
            if ref is not None or bigm.parent_fn is not None:
                have += 1
        dist.broadcast(have, src=owner)
        if int(have.item()) == 0:
            raise Exception("Key does {0} not exist, and no parent function prescripted".format(bigm.__shard_idx_to_key__(bidx)))
        if owner == grid.rank:
            tile = bigm.get_block(*bidx).reshape(shape).contiguous()
        else:
            tile = torch.empty(shape, dtype=bigm.torch_dtype, device=dev)
        dist.broadcast(tile, src=owner)
        tout[sl].copy_(tile.cpu())
    return out


def allreduce_max_int(value: int, device) -> int:
    import torch.distributed as dist
    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t.item())
