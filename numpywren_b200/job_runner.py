"""The LambdaPACK task loop on a B200: ``lambdapack_run(program, ...)``.

Replaces reference numpywren/job_runner.py (worker loop :316-370, read/compute/write pipeline
:224-309, ``LambdaPackExecutor.run`` :66-159).  The reference worker pulls one ready node from SQS,
GETs its tiles from S3, runs one NumPy kernel, PUTs the result, then bumps Redis edge counters.
Here one host thread walks the same ready queue and turns every node into asynchronous GPU work:

  * tiles are torch tensors in HBM taken *by reference* from the BigMatrix store (no copy, no
    serialisation); results are adopted by reference;
  * each node's kernel is enqueued on one of a pool of CUDA streams (high-priority streams for the
    critical-path ops chol / trsm / qr_factor — the reference's ``num_priorities`` idea,
    lambdapack.py:482) and cross-stream tile dependencies are CUDA events, so the host never waits
    for the GPU until the ready queue is empty;
  * SSA intermediates whose only reader is the node that replaces them (``S[i,j,k] → S[i+1,j,k]``,
    ``S[i,j,i] → O[j,i]``) are updated in place: the SSA version axis of ``S`` aliases one buffer
    (SURVEY §7 "SSA storage blow-up"); the single-reader property is read off the expanded DAG;
  * ``chol`` hands the inverted diagonal blocks of its factor to the ``trsm`` nodes that read it.

Program state (node status, edge sums, terminator count, counters) still goes through
``LambdaPackProgram.post_op`` so the reference's bookkeeping semantics are preserved.
"""
from __future__ import annotations

import os
import threading
import time
import traceback
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from . import kernels
from . import lambdapack as lp
from .compiler import ExpandedNode, _tile_key

HIGH_PRIORITY_KERNELS = ("chol", "trsm", "qr_factor", "qr_factor_triangular", "lq_factor")


class LRUCache(object):
    """Tile cache with least-recently-used eviction (reference job_runner.py:34-64).  Kept for the
    instruction-level API (RemoteRead.cache); the engine itself reads tiles in place from HBM."""

    def __init__(self, max_items=10):
        self.cache = {}
        self.key_order = []
        self.max_items = max_items

    def __setitem__(self, key, value):
        self.cache[key] = value
        self._mark(key)

    def __getitem__(self, key):
        value = self.cache[key]
        self._mark(key)
        return value

    def __contains__(self, obj):
        return obj in self.cache

    def _mark(self, key):
        if key in self.key_order:
            self.key_order.remove(key)
        self.key_order.insert(0, key)
        while len(self.key_order) > self.max_items:
            old = self.key_order.pop()
            del self.cache[old]


def calculate_busy_time(rtimes):
    """Union of [start, end] intervals (reference job_runner.py:162-175)."""
    events = sorted([(s, 1) for s, _ in rtimes] + [(e, -1) for _, e in rtimes])
    running, out, start = 0, [], 0
    for t, d in events:
        if running == 0 and d == 1:
            start = t
        if running == 1 and d == -1:
            out.append([start, t])
        running += d
    return out


class StreamPool:
    """Per-device CUDA streams the engine issues onto."""

    _pools: Dict[int, "StreamPool"] = {}

    def __init__(self, device: torch.device, n_normal: int, n_high: int):
        self.device = device
        with torch.cuda.device(device):
            self.normal = [torch.cuda.Stream(device=device, priority=0) for _ in range(n_normal)]
            self.high = [torch.cuda.Stream(device=device, priority=-1) for _ in range(n_high)]

    @classmethod
    def get(cls, device: torch.device, n_normal: int, n_high: int) -> "StreamPool":
        key = (device.index if device.index is not None else torch.cuda.current_device(), n_normal, n_high)
        if key not in cls._pools:
            cls._pools[key] = StreamPool(device, n_normal, n_high)
        return cls._pools[key]


class TileEngine:
    """Executes ready nodes of one LambdaPackProgram on one GPU (one process = one GPU)."""

    def __init__(self, program: lp.LambdaPackProgram, streams: int = 4, high_streams: int = 2, inplace: bool = True,
                 consume_inputs: bool = False, profile: bool = False, comm=None, free_intermediates: bool = False):
        self.program = program
        self.compiled = program.program
        self.inplace = inplace
        self.consume_inputs = consume_inputs
        self.profile = profile
        self.comm = comm
        self.device = None
        self.pool: Optional[StreamPool] = None
        self.n_streams, self.n_high = streams, high_streams
        self.tile_event: Dict[Any, Tuple[torch.cuda.Event, torch.cuda.Stream]] = {}
        self.invdiag: Dict[Any, torch.Tensor] = {}
        self.infos: List[Tuple[ExpandedNode, torch.Tensor]] = []
        self.timeline: List[Tuple[ExpandedNode, torch.cuda.Event, torch.cuda.Event, int]] = []
        self.launched = 0
        self._issue_lock = threading.RLock()
        self.input_names = set(self.compiled.inputs)
        self._input_mats = {id(self.compiled.scope[n]) for n in self.input_names if n in self.compiled.scope}
        self._prio: Optional[List[int]] = None
        self._lower_ok: Dict[int, bool] = {}
        self.skip_upper = os.environ.get("NPW_B200_SKIP_UPPER", "1") != "0"
        # dead-tile reclamation (opt-in): an SSA intermediate is dropped from the HBM store as soon as its last reader
        # has been enqueued.  The reference keeps every version in S3 for ever; in HBM the QR / BDFAC / GEMM programs'
        # intermediates would otherwise outgrow the device (GEMM Temp: M*N*K tiles).  On several GPUs a rank counts the
        # readers it executes itself: copies to other ranks are enqueued right after the producing node and keep their
        # own reference to the buffer (record_stream / pending send), so they do not pin the store entry.
        self.free_intermediates = free_intermediates
        self._keep_mats = {id(self.compiled.scope[n]) for n in list(self.compiled.inputs) + list(self.compiled.outputs)
                           if n in self.compiled.scope}
        self._reads_left: Dict[Any, int] = {}
        self.freed_tiles = 0
        # optional, off by default (NPW_B200_SYRK=i8emu): syrk products on the int8 tensor cores.
        # The int8 digits of a panel tile are a per-tile by-product like invdiag: extracted once, reused by every syrk
        # of the tile's block row / column, dropped after the last one.
        self.syrk_mode = os.environ.get("NPW_B200_SYRK", "native")
        self.i8_digits = int(os.environ.get("NPW_B200_I8_DIGITS", "6"))
        self._digits: Dict[Any, Tuple[Any, ...]] = {}     # tile key -> (digits, exponents, ready event, stream)
        self._digit_uses: Dict[Any, int] = {}

    # ------------------------------------------------------------------ priorities
    def priorities(self) -> List[int]:
        """Longest path to a sink (unit costs): the critical-path priority used to order the ready queue."""
        if self._prio is None:
            nodes = self.compiled.nodes
            prio = [0] * len(nodes)
            indeg = [len(n.children) for n in nodes]
            stack = [n.nid for n in nodes if not n.children]
            while stack:
                v = stack.pop()
                for p in nodes[v].parents:
                    if prio[v] + 1 > prio[p]:
                        prio[p] = prio[v] + 1
                    indeg[p] -= 1
                    if indeg[p] == 0:
                        stack.append(p)
            self._prio = prio
        return self._prio

    def _only_feeds_lower_readers(self, node) -> bool:
        """True when every consumer of this node's output is chol or another same-operand (diagonal) syrk."""
        ok = self._lower_ok.get(node.nid)
        if ok is None:
            ok = True
            for c in node.children:
                ch = self.compiled.nodes[c]
                if ch.call.compute is kernels.chol:
                    continue
                if ch.call.compute is kernels.syrk and len(ch.reads) == 3 and \
                        _tile_key(*ch.reads[1]) == _tile_key(*ch.reads[2]) and \
                        _tile_key(*ch.reads[0]) == _tile_key(*node.writes[0]) and self._only_feeds_lower_readers(ch):
                    continue  # the next diagonal update also ignores (and never exposes) the upper triangle
                ok = False
                break
            if not node.children:
                ok = False  # a sink's tile is user-visible: compute all of it
            self._lower_ok[node.nid] = ok
        return ok

    # ------------------------------------------------------------------ plumbing
    def _ensure_device(self, tensor_device: torch.device):
        if self.pool is None:
            if tensor_device.type != "cuda":
                raise kernels._capi.NpwError(
                    f"lambdapack_run: tiles live on {tensor_device}; the B200 engine has no CPU execution path")
            self.device = tensor_device
            self.pool = StreamPool.get(tensor_device, self.n_streams, self.n_high)
            # everything already in the store was produced on whatever stream the caller used
            self._entry_event = torch.cuda.Event()
            self._entry_event.record(torch.cuda.current_stream(tensor_device))
            self._entered = set()

    def _stream_for(self, node: ExpandedNode) -> torch.cuda.Stream:
        m, idx = node.writes[0]
        true_idx = m.true_block_idx(*idx)
        tail = true_idx[-2:] if len(true_idx) == 3 else true_idx
        h = 0
        for v in tail:
            h = h * 1000003 + int(v) + 7
        pool = self.pool.high if (node.call.compute_name in HIGH_PRIORITY_KERNELS and self.pool.high) else self.pool.normal
        return pool[h % len(pool)]

    def _wait_tile(self, key, stream):
        ev = self.tile_event.get(key)
        if ev is not None and ev[1] is not stream:
            stream.wait_event(ev[0])

    def _owned_input(self, m, idx, ref) -> bool:
        """May the node overwrite this stored tile?  Only if it is the tile's single reader and the tile is an
        intermediate (or inputs were declared consumable)."""
        if not self.inplace or ref is None:
            return False
        if id(m) in self._input_mats and not self.consume_inputs:
            return False
        if getattr(m, "transposed", False):
            return False
        return self.compiled.num_readers(m, idx) == 1

    def _read(self, m, idx, stream):
        """→ (2-D/N-D tile tensor, stored_ref_or_None, tile_key).  Never copies unless lambdav/transposed views force it."""
        key = _tile_key(m, idx)
        if self.comm is not None and self.comm.grid.owner(m, idx) != self.comm.rank:
            buf = self.comm.remote_tile(key, stream)
            if buf is None:
                raise Exception("tile {0}{1} lives on rank {2} and was never received".format(
                    m.key, list(idx), self.comm.grid.owner(m, idx)))
            return (buf.squeeze() if m.autosqueeze else buf), None, key, False
        self._wait_tile(key, stream)
        ref = m._get_block_ref(*idx) if hasattr(m, "_get_block_ref") else None
        if ref is not None and key not in self.tile_event:
            up = m._ready_event(*m.true_block_idx(*idx)) if hasattr(m, "_ready_event") else None
            if up is not None:             # tile still arriving from the host on the upload stream
                stream.wait_event(up)
                ref.record_stream(stream)
        shifted = (len(set(idx)) == 1 and len(set(m.shape)) == 1 and len(m.shape) != 1 and m.lambdav != 0)
        if ref is None or shifted:
            # default tile (parent_fn), transposed view, or diagonal shift: the public path makes a private copy
            tile = m.get_block(*idx)
            return tile, None, key, True
        if self.free_intermediates and id(m) not in self._keep_mats:
            ref.record_stream(stream)      # the allocator must not recycle the buffer before this stream is done with it
        tile = ref.squeeze() if m.autosqueeze else ref
        return tile, ref, key, False

    def _store(self, m, idx, tile, stream, event):
        shape = m.block_shape(*idx) if hasattr(m, "block_shape") else tuple(tile.shape)
        if tuple(tile.shape) != tuple(shape):
            if m.autosqueeze and list(tile.shape) == [x for x in shape if x != 1]:
                tile = tile.reshape(shape)
            elif m.safe:
                raise Exception("{2} Incompatible block size: {0} vs {1}".format(tuple(tile.shape), shape, m))
        if not tile.is_contiguous():
            tile = tile.contiguous()
        m._put_block_ref(tile, *idx)
        self.tile_event[_tile_key(m, idx)] = (event, stream)

    # ------------------------------------------------------------------ one node
    def run_node(self, node: ExpandedNode):
        """Enqueue one tile task.  Several runner threads may share the engine (the reference's tests start several
        workers per program): issuing is serialised here — it is host work only, the kernels run asynchronously."""
        with self._issue_lock:
            self._run_node(node)

    def _run_node(self, node: ExpandedNode):
        first_m = node.reads[0][0] if node.reads else node.writes[0][0]
        self._ensure_device(first_m.device)
        stream = self._stream_for(node)
        name = node.call.compute_name
        fn = node.call.compute
        with torch.cuda.stream(stream):
            if id(stream) not in self._entered:
                self._entered.add(id(stream))
                stream.wait_event(self._entry_event)
            t0 = None
            if self.profile:
                t0 = torch.cuda.Event(enable_timing=True)
                t0.record(stream)
            tiles, refs, keys, private = [], [], [], []
            for (m, idx) in node.reads:
                tile, ref, key, priv = self._read(m, idx, stream)
                tiles.append(tile); refs.append(ref); keys.append(key); private.append(priv)
            args = []
            for kind, j in node.arg_layout:
                if kind == "read":
                    args.append(tiles[j])
                elif isinstance(node.scalars[j], float):  # ints are dropped (reference lambdapack.py:364-368)
                    args.append(node.scalars[j])
            consumed = None  # index of the read whose buffer becomes the output

            def can_overwrite(j):
                m, idx = node.reads[j]
                return tiles[j].dim() == 2 and tiles[j].is_contiguous() and (
                    private[j] or self._owned_input(m, idx, refs[j]))

            if fn is kernels.syrk and len(args) == 3:
                out = args[0] if can_overwrite(0) else None
                consumed = 0 if out is not None else None
                # a diagonal tile S[i,j,j] = S - L_j L_j^T feeds only chol (lower triangle) or the next diagonal
                # syrk: the CTA tiles strictly above the diagonal are skipped (half the flops of these tasks)
                diag = self.skip_upper and keys[1] == keys[2] and self._only_feeds_lower_readers(node)
                if self.syrk_mode == "i8emu" and self._i8emu_ok(args):
                    xd, xe = self._tile_digits(node, 1, args[1], keys[1], stream)
                    yd, ye = self._tile_digits(node, 2, args[2], keys[2], stream)
                    results = kernels.syrk_i8emu(args[0], xd, xe, yd, ye, out=out, lower=diag)
                    self._release_digits(keys[1])
                    self._release_digits(keys[2])
                else:
                    results = kernels.syrk(args[0], args[1], args[2], out=out, lower=diag)
            elif fn is kernels.trsm and len(args) == 2:
                out = args[1] if can_overwrite(1) else None
                consumed = 1 if out is not None else None
                results = kernels.trsm_with_inverse(args[0], args[1], self.invdiag.get(keys[0]), out=out)
            elif fn is kernels.gemm_acc and len(args) == 3:
                out = args[0] if can_overwrite(0) else None
                consumed = 0 if out is not None else None
                results = kernels.gemm_acc(args[0], args[1], args[2], out=out)
            elif fn is kernels.chol and len(args) == 1:
                out = args[0] if can_overwrite(0) else None
                consumed = 0 if out is not None else None
                # trsm substitutes against the diagonal blocks of L itself (panel_kernel): no inverses to hand over
                L, info, _ = kernels.chol_async(args[0], want_inverse=False, out=out)
                self.infos.append((node, info))
                results = L
            else:
                results = fn(*args, **(node.call.kwargs or {}))
            if isinstance(results, tuple):
                if len(results) != len(node.writes):
                    raise Exception("Expected {0} results, got {1}".format(len(node.writes), len(results)))
            else:
                if len(node.writes) != 1:
                    raise Exception("Expected {0} results, got {1}".format(len(node.writes), 1))
                results = (results,)
            if consumed is not None and not private[consumed]:
                m, idx = node.reads[consumed]
                m.delete_block(*idx)  # the buffer now belongs to the output tile
            ev = torch.cuda.Event()
            for (m, idx), tile in zip(node.writes, results):
                self._store(m, idx, tile, stream, ev)
            ev.record(stream)
            for (m, idx), tile in zip(node.writes, results):
                if hasattr(m, "_after_put"):
                    m._after_put(m.true_block_idx(*idx), m._get_block_ref(*idx), ev)   # write-through host mirror
            if self.profile:
                t1 = torch.cuda.Event(enable_timing=True)
                t1.record(stream)
                self.timeline.append((node, t0, t1, id(stream)))
            if self.free_intermediates:
                self._release_dead_inputs(node, refs, keys, consumed)
        self.launched += 1
        # counters (reference job_runner.py:236-237, 265-266, 293-296)
        prog = self.program
        fl = getattr(fn, "flops", None)
        if fl is not None:
            try:
                prog.incr_flops(int(fl(*[a for a in args if isinstance(a, torch.Tensor)])))
            except Exception:
                pass
        for (m, _) in node.reads:
            prog.incr_read(lp._nbytes(m))
        for (m, _) in node.writes:
            prog.incr_write(lp._nbytes(m))

    # ------------------------------------------------------------------ experimental int8 emulation plumbing
    @staticmethod
    def _i8emu_ok(args) -> bool:
        s, x, y = args[0], args[1], args[2]
        return (x.dim() == 2 and y.dim() == 2 and x.shape[0] % 128 == 0 and y.shape[0] % 64 == 0 and x.shape[1] % 128 == 0
                and x.shape[1] == y.shape[1] and x.is_contiguous() and y.is_contiguous())

    def _syrk_uses(self, key) -> int:
        """How many syrk operand slots read this tile in the whole program (each needs the digits once)."""
        n = self._digit_uses.get(key)
        if n is None:
            n = 0
            for nid in set(self.compiled._readers.get(key, ())):   # a diagonal update lists the tile twice
                nd = self.compiled.nodes[nid]
                if nd.call.compute is kernels.syrk and len(nd.reads) == 3:
                    n += sum(1 for j in (1, 2) if _tile_key(*nd.reads[j]) == key)
            self._digit_uses[key] = n
        return n

    def _tile_digits(self, node, j, tile, key, stream):
        ent = self._digits.get(key)
        if ent is None:
            self._syrk_uses(key)
            d, e = kernels.split_i8(tile, self.i8_digits)
            ev = torch.cuda.Event()
            ev.record(stream)
            ent = self._digits[key] = (d, e, ev, stream)
        elif ent[3] is not stream:
            stream.wait_event(ent[2])
            ent[0].record_stream(stream)
            ent[1].record_stream(stream)
        return ent[0], ent[1]

    def _release_digits(self, key):
        left = self._digit_uses.get(key, 0) - 1
        self._digit_uses[key] = left
        if left <= 0:
            self._digits.pop(key, None)

    def _release_dead_inputs(self, node, refs, keys, consumed):
        """Drop stored intermediates whose every reader has now been enqueued (stream order + record_stream keep the
        memory alive until the kernels that were given the pointer have run)."""
        written = {_tile_key(m, idx) for m, idx in node.writes}
        for j, (m, idx) in enumerate(node.reads):
            if refs[j] is None or id(m) in self._keep_mats or j == consumed:
                continue
            key = keys[j]
            left = self._reads_left.get(key)
            if left is None:
                if self.comm is None:
                    left = self.compiled.num_readers(m, idx)
                else:
                    mine = self.comm.rank
                    left = sum(1 for r in self.compiled._readers.get(key, ()) if self.comm.plan.exec_rank[r] == mine)
            left -= 1
            self._reads_left[key] = left
            if left <= 0 and key not in written:
                m.delete_block(*idx)
                self.tile_event.pop(key, None)
                self.freed_tiles += 1

    # ------------------------------------------------------------------ drain
    def finish(self):
        """Wait for the device and surface asynchronous kernel failures (LAPACK-style info codes)."""
        if self.comm is not None:
            self.comm.drain()
        if self.device is not None:
            torch.cuda.synchronize(self.device)
        bad = []
        with self._issue_lock:
            infos, self.infos = self.infos, []
        for node, info in infos:
            code = int(info.item())
            if code != 0:
                bad.append((node, code))
        if self.comm is not None:
            # every rank must agree on failure: share the smallest failing node id (or "none")
            from . import parallel
            mine = min((n.nid for n, _ in bad), default=-1)
            worst = parallel.allreduce_max_int(-1 if mine < 0 else (1 << 40) - mine, self.device or torch.device("cuda"))
            if worst >= 0 and not bad:
                nid = (1 << 40) - worst
                bad = [(self.compiled.nodes[nid], -1)]
        return bad


def _engine_for(program: lp.LambdaPackProgram, **opts) -> TileEngine:
    eng = getattr(program, "_engine", None)
    if eng is None:
        from . import parallel
        grid = parallel.current_grid()
        if grid is not None and grid.world > 1:
            opts["comm"] = parallel.make_exchange(program.program, grid, torch.device("cuda", torch.cuda.current_device()))
        eng = TileEngine(program, **opts)
        program._engine = eng
        prio = eng.priorities()
        compiled = program.program
        program._priority_fn = lambda e, v: prio[compiled.node(e, v).nid]
        program._prio_by_nid = prio
    return eng


def _engine_options(pipeline_width=5, streams=None, high_streams=None, inplace=None, consume_inputs=False, profile=False,
                    free_intermediates=None):
    """Resolve the engine's tunables from keywords and NPW_B200_* environment variables."""
    if streams is not None:
        n_streams = streams
    elif "NPW_B200_STREAMS" in os.environ:
        n_streams = int(os.environ["NPW_B200_STREAMS"])
    else:
        from . import parallel
        grid = parallel.current_grid()
        # several GPUs: more streams, so that a task waiting for a panel tile from another GPU stalls fewer local ones
        n_streams = 8 if (grid is not None and grid.world > 1) else max(1, min(8, pipeline_width))
    n_high = high_streams if high_streams is not None else int(os.environ.get("NPW_B200_HIGH_STREAMS", 2))
    if inplace is None:
        inplace = os.environ.get("NPW_B200_INPLACE", "1") != "0"
    if free_intermediates is None:
        free_intermediates = os.environ.get("NPW_B200_FREE_INTERMEDIATES", "0") != "0"
    return dict(streams=n_streams, high_streams=n_high, inplace=inplace, consume_inputs=consume_inputs, profile=profile,
                free_intermediates=free_intermediates)


def prepare(program, pipeline_width=5, **engine_kwargs):
    """Build everything about ``program`` that depends only on the DAG, ahead of ``lambdapack_run``: the expanded node
    list, the critical-path priorities and, on several GPUs, the transfer plan with its inbox slots (and the symmetric
    inbox itself, a collective allocation).  Optional — ``lambdapack_run`` does the same on first use — but it keeps
    this static analysis (0.2 s of Python for the 5984-node Cholesky) out of the time the GPUs spend on the program,
    like the DAG expansion itself.  Takes the engine keywords of ``lambdapack_run``; returns seconds spent."""
    t0 = time.time()
    program.program.nodes
    _engine_for(program, **_engine_options(pipeline_width, **engine_kwargs))
    return time.time() - t0


def lambdapack_run(program, pipeline_width=5, msg_vis_timeout=60, cache_size=5, timeout=200, idle_timeout=5,
                   msg_vis_timeout_jitter=15, compute_threads=1, streams=None, high_streams=None, inplace=None,
                   consume_inputs=False, profile=False, free_intermediates=None):
    """Run ready nodes of ``program`` until it finishes, fails, or ``timeout`` seconds elapse.

    Signature and return keys follow reference job_runner.lambdapack_run (:316-370).  ``pipeline_width``
    (the reference's read/compute/write overlap depth) sets the number of CUDA streams unless ``streams``
    is given; ``cache_size``, ``msg_vis_timeout*`` and ``compute_threads`` have no meaning without
    S3/SQS/BLAS threads and are accepted for compatibility.  Extra keywords tune the B200 engine (they are ignored when
    ``prepare`` already built the engine for this program).
    """
    program.incr_up(1)
    with program._lock:
        program._runner_active += 1
    lambda_start = time.time()
    eng = _engine_for(program, **_engine_options(pipeline_width, streams, high_streams, inplace, consume_inputs, profile,
                                                 free_intermediates))
    program._defer_success = True
    nodes = program.program.nodes
    executed, refs = [], []
    if eng.comm is not None:
        with program._lock:
            if program._runner_active > 1:
                # the tile exchange pairs the k-th send with the k-th receive of every (src, dst): all ranks must walk
                # the DAG in the same order, which several runner threads on one rank cannot guarantee
                program._runner_active -= 1
                program.decr_up(1)
                raise RuntimeError("lambdapack_run: one runner thread per rank when the program is sharded over GPUs")
    counted = True
    next_node = None      # eager=True: post_op hands one ready child straight back to this runner (reference :113-139)
    try:
        while program.program_status() == lp.PS.RUNNING:
            if time.time() - lambda_start > timeout:
                if next_node is not None:      # not run: back to the queue, it is READY
                    prio = getattr(program, "_prio_by_nid", None)
                    program._enqueue_node(next_node, prio[next_node.nid] if prio is not None else 0)
                    next_node = None
                break
            if next_node is not None:
                node, next_node = next_node, None
                expr_idx = node.expr_idx
            else:
                item = program._dequeue_item()
                if item is None:
                    break
                expr_idx, frozen, nid = item
                node = nodes[nid] if nid is not None else program.program.node(expr_idx, dict(frozen))
            var_values = node.var_values
            status = program.node_status_of(node)
            if status == lp.NS.FINISHED:
                program.incr_repeated_finish()
                continue
            if status == lp.NS.NOT_READY:
                program.incr_not_ready()
                continue
            try:
                if status in (lp.NS.READY, lp.NS.RUNNING):
                    if status == lp.NS.RUNNING:
                        program.incr_repeated_compute()
                    program.set_node_status_of(node, lp.NS.RUNNING)
                    comm = eng.comm
                    if comm is not None:
                        comm.before_node(node, eng)
                    if comm is None or comm.exec_rank(node) == comm.rank:
                        eng.run_node(node)
                    if comm is not None:
                        comm.after_node(node, eng)
                else:
                    program.incr_repeated_post_op()
                nxt, _ = program.post_op_node(node, lp.PS.SUCCESS)
                program.set_node_status_of(node, lp.NS.FINISHED)
                if nxt is not None:
                    next_node = program.program.node(nxt[0], nxt[1])
            except Exception:
                tb = traceback.format_exc()
                program.handle_exception("EXCEPTION", tb=tb, expr_idx=expr_idx, var_values=var_values)
                raise
            executed.append((expr_idx, var_values))
            refs.append((expr_idx, var_values))
        bad = eng.finish()
        if bad:
            node, code = bad[0]
            program.handle_exception("COMPUTE EXCEPTION", tb="", expr_idx=node.expr_idx, var_values=node.var_values)
            raise np.linalg.LinAlgError(
                "Matrix is not positive definite (tile task {0}{1}: leading minor {2})".format(
                    node.call.compute_name, node.var_values, code))
        # Only the LAST runner to leave may publish SUCCESS: its finish() (device drained, chol info codes checked)
        # happens after every other runner's last issue, because a runner leaves the count only here, after its own
        # finish().  A runner that drained early and then sees the flag set by a peer must not publish for it.
        with program._lock:
            program._runner_active -= 1
            counted = False
            if program._runner_active == 0 and getattr(program, "_all_terminators_done", False) \
                    and program._status == lp.PS.RUNNING:
                program._status = lp.PS.SUCCESS
    finally:
        program.decr_up(1)
        if counted:
            with program._lock:
                program._runner_active -= 1
    lambda_stop = time.time()
    return {"up_time": [lambda_start, lambda_stop],
            "exec_time": calculate_busy_time([[lambda_start, lambda_stop]]) if executed else [],
            "executed_messages": executed,
            "operator_refs": refs,
            "log": None}


def node_timeline(program):
    """Per-node GPU times (ms) recorded when ``profile=True``: [(compute_name, var_values, start_ms, end_ms, stream)]."""
    eng = getattr(program, "_engine", None)
    if eng is None or not eng.timeline:
        return []
    base = eng.timeline[0][1]
    out = []
    for node, t0, t1, sid in eng.timeline:
        out.append((node.call.compute_name, dict(node.var_values), base.elapsed_time(t0), base.elapsed_time(t1), sid))
    return out
