"""LambdaPACK runtime objects: instructions and program state.

Mirrors reference numpywren/lambdapack.py: the per-task instruction triple ``RemoteRead`` /
``RemoteCall`` / ``RemoteWrite`` inside an ``InstructionBlock`` (:206-474) and ``LambdaPackProgram``
(:477-777) with its node-status machine, edge counters, terminator counter, progress counters and
``start / wait / stop / free / post_op``.

What changed underneath: the reference keeps this state in Redis (atomic WATCH/MULTI increments,
:71-196) and the ready queue in SQS because its workers are stateless Lambdas.  Here the workers are
CUDA streams driven from one host process, so node status, edge sums and the ready queue are plain
in-process structures guarded by a lock; the semantics (a child becomes READY when its edge sum
equals its parent count; the program SUCCEEDS when the terminator count reaches
``num_terminators``) are unchanged.
"""
from __future__ import annotations

import heapq
import threading
import time
import traceback
from enum import Enum

import numpy as np
import torch


class RemoteInstructionOpCodes(Enum):
    S3_LOAD = 0
    S3_WRITE = 1
    GENERIC = 3
    RET = 4


class NodeStatus(Enum):
    NOT_READY = 0
    READY = 1
    RUNNING = 2
    POST_OP = 3
    FINISHED = 4


class EdgeStatus(Enum):
    NOT_READY = 0
    READY = 1


class ProgramStatus(Enum):
    SUCCESS = 0
    RUNNING = 1
    EXCEPTION = 2
    NOT_STARTED = 3


OC = RemoteInstructionOpCodes
NS = NodeStatus
ES = EdgeStatus
PS = ProgramStatus


_NBYTES_CACHE = {}


def _nbytes(matrix):
    """Bytes of one full tile of ``matrix`` (the reference's read/write counters, lambdapack.py:234,293); cached per
    (shard sizes, dtype) — the engine asks once per tile read or written."""
    key = (tuple(matrix.shard_sizes), str(matrix.dtype))
    v = _NBYTES_CACHE.get(key)
    if v is None:
        v = _NBYTES_CACHE[key] = int(np.prod(matrix.shard_sizes)) * np.dtype(matrix.dtype).itemsize
    return v


class RemoteInstruction(object):
    def __init__(self, i_id):
        self.id = i_id
        self.ret_code = -1
        self.start_time = None
        self.end_time = None
        self.type = None
        self.executor = None
        self.cache = None
        self.run = False
        self.read_size = 0
        self.write_size = 0

    def get_flops(self):
        return 0

    def clear(self):
        self.result = None


class RemoteRead(RemoteInstruction):
    """Fetch one tile (reference :225-283).  ``cache`` may be an LRUCache keyed like the reference's
    ``(key, bucket, true_block_idx)``."""

    def __init__(self, i_id, matrix, *bidxs):
        super().__init__(i_id)
        self.i_code = OC.S3_LOAD
        self.matrix = matrix
        self.bidxs = bidxs
        self.result = None
        self.cache_hit = False
        self.read_size = _nbytes(matrix)

    def __call__(self):
        self.start_time = time.time()
        if self.result is None:
            cache_key = (self.matrix.key, self.matrix.bucket, self.matrix.true_block_idx(*self.bidxs))
            if self.cache is not None and cache_key in self.cache:
                self.result = self.cache[cache_key]
                self.cache_hit = True
            else:
                self.result = self.matrix.get_block(*self.bidxs)
                if self.cache is not None:
                    self.cache[cache_key] = self.result
            self.size = self.result.numel() * self.result.element_size()
        self.end_time = time.time()
        return self.result

    def __str__(self):
        return "{0} = S3_LOAD {1} {2} {3}".format(self.id, self.matrix, len(self.bidxs), " ".join(str(x) for x in self.bidxs))


class RemoteWrite(RemoteInstruction):
    """Store one result tile (reference :285-341)."""

    def __init__(self, i_id, matrix, data_loc, data_idx, *bidxs):
        super().__init__(i_id)
        self.i_code = OC.S3_WRITE
        self.matrix = matrix
        self.bidxs = bidxs
        self.data_loc = data_loc
        self.data_idx = data_idx
        self.result = None
        self.sparse_write = False
        self.write_size = _nbytes(matrix)

    def __call__(self, skip_empty=False):
        self.start_time = time.time()
        if self.result is None:
            data = self.data_loc[self.data_idx]
            cache_key = (self.matrix.key, self.matrix.bucket, self.matrix.true_block_idx(*self.bidxs))
            if self.cache is not None:
                self.cache[cache_key] = data
            if skip_empty and bool((data == 0).all()):
                self.sparse_write = True  # block_sparse: an all-zero tile is left to parent_fn (reference :311-317)
            else:
                self.matrix.put_block(data, *self.bidxs)
            self.size = data.numel() * data.element_size()
            self.ret_code = 0
            self.result = 0
        self.end_time = time.time()
        return self.result

    def clear(self):
        self.result = None
        self.data_loc = None

    def __str__(self):
        return "{0} = S3_WRITE {1} {2} {3} {4}".format(self.id, self.matrix, len(self.bidxs),
                                                         " ".join(str(x) for x in self.bidxs), self.data_idx)


class RemoteCall(RemoteInstruction):
    """Run the tile kernel on the results of the reads (reference :344-415).

    Arguments are the RemoteRead results in DSL order plus Python ``float`` literals; ``int``
    arguments are dropped, exactly like reference :364-368.
    """

    def __init__(self, i_id, compute, argv_instr, num_outputs, symbols, **kwargs):
        super().__init__(i_id)
        self.i_code = OC.GENERIC
        self.results = [None for _ in range(num_outputs)]
        self.kwargs = kwargs
        self.compute = compute
        self.symbols = symbols
        self.argv_instr = argv_instr

    def _pyargs(self):
        out = []
        for arg in self.argv_instr:
            if isinstance(arg, RemoteRead):
                out.append(arg.result)
            elif isinstance(arg, float):
                out.append(arg)
        return out

    def __call__(self, prev=None):
        self.start_time = time.time()
        results = self.compute(*self._pyargs(), **self.kwargs)
        if isinstance(results, tuple) and len(results) != len(self.results):
            raise Exception("Expected {0} results, got {1}".format(len(self.results), len(results)))
        elif isinstance(results, tuple):
            for i, r in enumerate(results):
                self.results[i] = r
        else:
            self.results[0] = results
        self.ret_code = 0
        self.end_time = time.time()
        return self.results

    def clear(self):
        self.results = [None for _ in self.results]
        self.argv_instr = [None for _ in self.argv_instr]

    def get_flops(self):
        fl = getattr(self.compute, "flops", None)
        if fl is None:
            return 0
        return fl(*self._pyargs())

    def __str__(self):
        outs = ",".join(str(i + len(self.symbols)) for i in range(len(self.results)))
        return "{1} = {0}({2}, **kwargs)".format(getattr(self.compute, "__name__", self.compute), outs, ",".join(self.symbols))


class RemoteReturn(RemoteInstruction):
    def __init__(self, i_id):
        super().__init__(i_id)
        self.i_code = OC.RET
        self.result = None

    def __call__(self, program):
        self.start_time = time.time()
        program.return_success()
        self.end_time = time.time()
        return self.result

    def __str__(self):
        return "RET"


class InstructionBlock(object):
    block_count = 0

    def __init__(self, instrs, label=None, priority=0):
        self.instrs = instrs
        self.label = label
        self.priority = priority
        if self.label is None:
            self.label = "%{0}".format(InstructionBlock.block_count)
        InstructionBlock.block_count += 1
        self.start_time = None
        self.end_time = None

    def __call__(self):
        return [x() for x in self.instrs]

    def __str__(self):
        return self.label + "\n" + "".join("\t{0}\n".format(i) for i in self.instrs)

    def clear(self):
        [x.clear() for x in self.instrs]

    def total_flops(self):
        return sum(getattr(x, "flops", 0) for x in self.instrs)

    def total_io(self):
        return sum(getattr(x, "size", 0) for x in self.instrs)

    def __copy__(self):
        return InstructionBlock(self.instrs.copy(), self.label)


class LambdaPackProgram(object):
    """A compiled program plus its global execution state (reference :477-777)."""

    _hash_lock = threading.Lock()
    _hash_count = 0

    def __init__(self, program, config=None, num_priorities=1, eager=False, block_sparse=False):
        self.config = config
        self.program = program
        self.block_sparse = block_sparse
        self.max_priority = num_priorities - 1
        self.eager = eager
        with LambdaPackProgram._hash_lock:
            LambdaPackProgram._hash_count += 1
            # the reference uses str(int(time.time())) (:495); a counter keeps concurrent programs distinct
            self.hash = "{0}_{1}".format(int(time.time()), LambdaPackProgram._hash_count)
        self.up = 'up' + self.hash
        self._lock = threading.RLock()
        self._status = PS.NOT_STARTED
        self._node_status = {}
        self._edge_sum = {}
        self._edges_seen = set()
        self._terminators_done = set()
        self._ready = []          # heap of (-priority, seq, expr_idx, frozen var_values)
        self._seq = 0
        self._counters = {}
        self._exceptions = []
        self._runner_active = 0
        self.profiles = {}
        self.set_up(0)

    # ------------------------------------------------------------------ keys (reference :507-519)
    def _node_str(self, expr_idx, var_values):
        var_strs = sorted(["{0}:{1}".format(key, value) for key, value in var_values.items()])
        return "{0}_({1})".format(expr_idx, "-".join(var_strs))

    def _node_key(self, expr_idx, var_values):
        return "{0}_{1}".format(self.hash, self._node_str(expr_idx, var_values))

    def _node_edge_sum_key(self, expr_idx, var_values):
        return "{0}_{1}_edgesum".format(self.hash, self._node_str(expr_idx, var_values))

    def _edge_key(self, expr_idx1, var_values1, expr_idx2, var_values2):
        return "{0}_{1}_{2}".format(self.hash, self._node_str(expr_idx1, var_values1),
                                    self._node_str(expr_idx2, var_values2))

    def _node_str_of(self, node):
        """_node_str of an expanded node, formatted once (the reference re-formats these Redis keys on every access;
        with ~15 accesses per node that was half of the host loop's time)."""
        cache = self.__dict__.setdefault("_nstr_cache", {})
        s = cache.get(node.nid)
        if s is None:
            s = cache[node.nid] = self._node_str(node.expr_idx, node.var_values)
        return s

    def node_status_of(self, node):
        with self._lock:
            return self._node_status.get(self._node_str_of(node), NS.NOT_READY)

    def set_node_status_of(self, node, status):
        with self._lock:
            self._node_status[self._node_str_of(node)] = status
        return status

    # ------------------------------------------------------------------ node / program status
    def get_node_status(self, expr_idx, var_values):
        with self._lock:
            return self._node_status.get(self._node_str(expr_idx, var_values), NS.NOT_READY)

    def set_node_status(self, expr_id, var_values, status):
        with self._lock:
            self._node_status[self._node_str(expr_id, var_values)] = status
        return status

    def program_status(self):
        with self._lock:
            return self._status

    def return_success(self):
        with self._lock:
            if self._status != PS.EXCEPTION:
                self._status = PS.SUCCESS

    def stop(self):
        with self._lock:
            self._exceptions.append("EXCEPTION.DRIVER.CANCELLED: cancelled by driver")
            self._status = PS.EXCEPTION

    def handle_exception(self, error, tb, expr_idx, var_values):
        with self._lock:
            self._exceptions.append("EXCEPTION.{0}: {1}{2}".format(self._node_str(expr_idx, var_values), tb, error))
            self._status = PS.EXCEPTION

    @property
    def exceptions(self):
        return list(self._exceptions)

    # ------------------------------------------------------------------ ready queue (replaces SQS)
    def _enqueue(self, expr_idx, var_values, priority=None):
        if priority is None:
            fn = getattr(self, "_priority_fn", None)
            priority = fn(expr_idx, var_values) if fn is not None else 0
        with self._lock:
            self._seq += 1
            heapq.heappush(self._ready, (-priority, self._seq, int(expr_idx), tuple(sorted(var_values.items())), None))

    def _enqueue_node(self, node, priority):
        """Same queue, for a node of the expanded DAG (no key formatting: the host loop's fast path)."""
        with self._lock:
            self._seq += 1
            heapq.heappush(self._ready, (-priority, self._seq, node.expr_idx, node.key[1], node.nid))

    def _dequeue(self):
        with self._lock:
            if not self._ready:
                return None
            _, _, e, v, _ = heapq.heappop(self._ready)
            return e, dict(v)

    def _dequeue_item(self):
        """-> (expr_idx, frozen var_values, node id or None) of the highest-priority ready node, or None."""
        with self._lock:
            if not self._ready:
                return None
            _, _, e, v, nid = heapq.heappop(self._ready)
            return e, v, nid

    def queue_depth(self):
        with self._lock:
            return len(self._ready)

    def start(self, parallel=False):
        """Mark the program RUNNING and make every starter READY (reference :641-658)."""
        with self._lock:
            self._status = PS.RUNNING
            for x in self.program.starters:
                self.set_node_status(*x, NS.READY)
                self._enqueue(x[0], x[1])
        return 0

    def conditional_increment(self, child, edge):
        """Count edge → child at most once and return the child's edge sum (reference :154-196)."""
        with self._lock:
            if edge not in self._edges_seen:
                self._edges_seen.add(edge)
                self._edge_sum[child] = self._edge_sum.get(child, 0) + 1
            return self._edge_sum.get(child, 0)

    def post_op(self, expr_idx, var_values, ret_code, inst_block, tb=None):
        """After a node's writes landed: release its children, count terminators (reference :545-639)."""
        try:
            post_op_start = time.time()
            children = self.program.find_children(expr_idx, var_values)
            self.set_node_status(expr_idx, var_values, NS.POST_OP)
            if ret_code == PS.EXCEPTION and tb is not None:
                self.handle_exception(" EXCEPTION", tb=tb, expr_idx=expr_idx, var_values=var_values)
            ready_children = []
            for child in children:
                my_child_edge = self._edge_key(expr_idx, var_values, *child)
                val = self.conditional_increment(self._node_edge_sum_key(*child), my_child_edge)
                num_child_parents = len(self.program.find_parents(child[0], child[1]))
                if val == num_child_parents and self.get_node_status(*child) != NS.FINISHED:
                    self.set_node_status(*child, NS.READY)
                    ready_children.append(child)
            next_operator = None
            if self.eager and ready_children:
                next_operator = ready_children.pop()
            for child in ready_children:
                self._enqueue(child[0], child[1])
            if inst_block is not None:
                inst_block.end_time = time.time()
                inst_block.clear()
                inst_block.post_op_start = post_op_start
                inst_block.post_op_end = time.time()
                inst_block.expr_idx = expr_idx
                inst_block.var_values = var_values
            self.incr_progress()
            if self.program.is_terminator(expr_idx):
                with self._lock:
                    self._terminators_done.add(self._node_str(expr_idx, var_values))
                    if len(self._terminators_done) == self.program.num_terminators:
                        # with an asynchronous engine "posted" is not "executed": the runner flips the
                        # status to SUCCESS after the device has drained (job_runner.lambdapack_run)
                        self._all_terminators_done = True
                        if not getattr(self, "_defer_success", False):
                            self.return_success()
            return next_operator, None
        except Exception as e:
            tb = traceback.format_exc()
            self.handle_exception("POST OP EXCEPTION", tb=tb, expr_idx=expr_idx, var_values=var_values)
            raise

    def post_op_node(self, node, ret_code):
        """``post_op`` for a node of the expanded DAG: identical state transitions and keys (edge sums, READY marks,
        terminator set — the public accessors see the same dictionaries), but children / parents come straight from the
        DAG arrays and every key string is formatted once per node instead of once per access."""
        try:
            nodes = self.program.nodes
            me = self._node_str_of(node)
            prio = getattr(self, "_prio_by_nid", None)
            ready_children = []
            with self._lock:
                self._node_status[me] = NS.POST_OP
                for c in node.children:
                    child = nodes[c]
                    cs = self._node_str_of(child)
                    edge = "{0}_{1}_{2}".format(self.hash, me, cs)
                    ckey = "{0}_{1}_edgesum".format(self.hash, cs)
                    if edge not in self._edges_seen:
                        self._edges_seen.add(edge)
                        self._edge_sum[ckey] = self._edge_sum.get(ckey, 0) + 1
                    if self._edge_sum.get(ckey, 0) == len(child.parents) and \
                            self._node_status.get(cs, NS.NOT_READY) != NS.FINISHED:
                        self._node_status[cs] = NS.READY
                        ready_children.append(child)
            next_operator = None
            if self.eager and ready_children:
                last = ready_children.pop()
                next_operator = last.ref
            for child in ready_children:
                if prio is not None:
                    self._enqueue_node(child, prio[child.nid])
                else:
                    self._enqueue(child.expr_idx, child.var_values)
            self.incr_progress()
            if self.program.is_terminator(node.expr_idx):
                with self._lock:
                    self._terminators_done.add(me)
                    if len(self._terminators_done) == self.program.num_terminators:
                        self._all_terminators_done = True
                        if not getattr(self, "_defer_success", False):
                            self.return_success()
            return next_operator, None
        except Exception as e:
            tb = traceback.format_exc()
            self.handle_exception("POST OP EXCEPTION", tb=tb, expr_idx=node.expr_idx, var_values=node.var_values)
            raise

    # ------------------------------------------------------------------ counters (reference :683-752)
    def _incr(self, name, amount=1):
        with self._lock:
            self._counters[name] = self._counters.get(name, 0) + amount

    def _get(self, name):
        with self._lock:
            return self._counters.get(name, 0)

    def incr_up(self, amount):
        self._incr("up", amount)

    def decr_up(self, amount):
        self._incr("up", -amount)

    def get_up(self):
        return self._get("up")

    def set_up(self, value):
        with self._lock:
            self._counters["up"] = value

    def incr_repeated_compute(self, amount=1):
        self._incr("repeated_compute", amount)

    def incr_repeated_post_op(self, amount=1):
        self._incr("repeated_post_op", amount)

    def incr_repeated_finish(self, amount=1):
        self._incr("repeated_finish", amount)

    def incr_not_ready(self, amount=1):
        self._incr("not_ready", amount)

    def incr_progress(self):
        self._incr("progress", 1)

    def incr_flops(self, amount):
        if amount > 0:
            self._incr("flops", amount)

    def incr_read(self, amount):
        if amount > 0:
            self._incr("read", amount)

    def incr_sparse_read(self, amount):
        if amount > 0:
            self._incr("sparse_read", amount)

    def incr_write(self, amount):
        if amount > 0:
            self._incr("write", amount)

    def incr_sparse_write(self, amount):
        if amount > 0:
            self._incr("write_sparse", amount)

    def decr_flops(self, amount):
        if amount > 0:
            self._incr("flops", -amount)

    def decr_read(self, amount):
        if amount > 0:
            self._incr("read", -amount)

    def decr_write(self, amount):
        if amount > 0:
            self._incr("write", -amount)

    def get_flops(self):
        return self._get("flops")

    def get_read(self):
        return self._get("read")

    def get_write(self):
        return self._get("write")

    def get_progress(self):
        return self._get("progress")

    # ------------------------------------------------------------------ driver side
    def wait(self, sleep_time=1):
        """Block until the program leaves RUNNING (reference :754-760).

        In the reference some other process runs ``job_runner.lambdapack_run``; here, if no runner is
        attached to this program, the calling thread becomes the runner (there is nothing else that
        could make progress), otherwise it polls.
        """
        while self.program_status() == PS.RUNNING:
            with self._lock:
                idle = self._runner_active == 0
            if idle:
                from . import job_runner
                job_runner.lambdapack_run(self)
                if self.program_status() == PS.RUNNING and self.queue_depth() == 0:
                    break  # nothing runnable and not finished: do not spin forever
            else:
                time.sleep(min(sleep_time, 0.01))
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    def free(self):
        with self._lock:
            self._ready = []

    def get_profiling_info(self, expr_idx, var_values):
        return self.profiles.get(self._node_str(expr_idx, var_values))

    def get_all_profiling_info(self):
        return list(self.profiles.values())
