"""LambdaPackProgram bookkeeping (node status machine, edge sums, terminators, counters) driven through the
instruction-level API on a storage-only host device.  Only ``identity`` nodes run, so no kernel is needed:
this is the reference's post_op logic (lambdapack.py:545-639) without Redis/SQS."""
import numpy as np
import pytest

from numpywren_b200 import algs, compiler, job_runner
from numpywren_b200 import lambdapack as lp
from numpywren_b200.matrix import BigMatrix
from numpywren_b200.matrix_init import shard_matrix


def Chain(A: BigMatrix, B: BigMatrix, C: BigMatrix, N: int):
    for i in range(N):
        B[i, 0] = identity(A[i, 0])
    for i in range(N):
        for j in range(0, 2):
            C[i, j] = identity(B[i, 0])


def build(unique_key, N=4):
    X = np.arange(N * 4, dtype=np.float64).reshape(2 * N, 2)
    A = BigMatrix(unique_key("A"), shape=X.shape, shard_sizes=(2, 2), device="cpu")
    B = BigMatrix(unique_key("B"), shape=X.shape, shard_sizes=(2, 2), device="cpu")
    C = BigMatrix(unique_key("C"), shape=(2 * N, 4), shard_sizes=(2, 2), device="cpu")
    for m in (A, B, C):
        m.free()
    shard_matrix(A, X)
    p = compiler.lpcompile_for_execution(Chain, inputs=["A"], outputs=["C"])(A, B, C, N)
    return X, A, C, lp.LambdaPackProgram(p)


def drain(program, cache=None):
    """A synchronous worker built from the public pieces: dequeue -> eval_expr -> instrs -> post_op
    (the status handling of reference job_runner.py:101-135)."""
    order = []
    while True:
        item = program._dequeue()
        if item is None:
            break
        e, v = item
        if program.get_node_status(e, v) == lp.NS.FINISHED:
            program.incr_repeated_finish()
            continue
        assert program.get_node_status(e, v) == lp.NS.READY
        program.set_node_status(e, v, lp.NS.RUNNING)
        ib = program.program.eval_expr(e, v)
        for ins in ib.instrs:
            ins.cache = cache
            ins()
        program.post_op(e, v, lp.PS.SUCCESS, ib)
        program.set_node_status(e, v, lp.NS.FINISHED)
        order.append((e, tuple(sorted(v.items()))))
    return order


def test_status_machine_and_result(unique_key):
    N = 4
    X, A, C, program = build(unique_key, N)
    assert program.program_status() == lp.PS.NOT_STARTED
    assert len(program.program.nodes) == N + 2 * N
    assert program.program.num_terminators == 2 * N
    program.start()
    assert program.program_status() == lp.PS.RUNNING
    assert program.queue_depth() == len(program.program.starters) == N
    order = drain(program, cache=job_runner.LRUCache(4))
    assert len(order) == 3 * N and len(set(order)) == 3 * N
    assert program.program_status() == lp.PS.SUCCESS
    assert program.get_progress() == 3 * N
    assert np.array_equal(C.numpy(), np.hstack([X, X]))
    # every first-phase node ran before the nodes that read its tile
    pos = {k: i for i, k in enumerate(order)}
    for i in range(N):
        for j in range(2):
            assert pos[(0, (("i", i),))] < pos[(1, (("i", i), ("j", j)))]


def test_children_become_ready_only_after_all_parents(unique_key):
    O = BigMatrix(unique_key("O"), shape=(8, 8), shard_sizes=(2, 2), device="cpu")
    I = BigMatrix(unique_key("I"), shape=(8, 8), shard_sizes=(2, 2), device="cpu")
    S = BigMatrix(unique_key("S"), shape=(5, 8, 8), shard_sizes=(1, 2, 2), device="cpu")
    p = compiler.lpcompile_for_execution(algs.CHOLESKY, ["I"], ["O"])(O, I, S, 4, 0)
    program = lp.LambdaPackProgram(p)
    program.start()
    assert program._dequeue() == (0, {})
    assert program.queue_depth() == 0
    program.post_op(0, {}, lp.PS.SUCCESS, None)
    # chol(0) releases the three trsm(j, 0); syrk nodes need two trsm parents each
    ready = sorted(program._dequeue()[1]["j"] for _ in range(3))
    assert ready == [1, 2, 3] and program.queue_depth() == 0
    program.post_op(1, {"j": 1}, lp.PS.SUCCESS, None)
    assert program.get_node_status(2, {"j": 1, "k": 1}) == lp.NS.READY          # needs only trsm(1)
    assert program.get_node_status(2, {"j": 2, "k": 1}) == lp.NS.NOT_READY      # needs trsm(1) and trsm(2)
    program.post_op(1, {"j": 1}, lp.PS.SUCCESS, None)                           # replayed post_op: edges count once
    assert program.get_node_status(2, {"j": 2, "k": 1}) == lp.NS.NOT_READY
    program.post_op(1, {"j": 2}, lp.PS.SUCCESS, None)
    assert program.get_node_status(2, {"j": 2, "k": 1}) == lp.NS.READY
    assert program.get_node_status(2, {"j": 2, "k": 2}) == lp.NS.READY
    assert program.program_status() == lp.PS.RUNNING


def test_counters_and_stop(unique_key):
    _, _, _, program = build(unique_key)
    program.incr_flops(10); program.incr_flops(-5); program.incr_read(7); program.incr_write(3)
    assert (program.get_flops(), program.get_read(), program.get_write()) == (10, 7, 3)
    program.decr_flops(4)
    assert program.get_flops() == 6
    program.incr_up(2); program.decr_up(1)
    assert program.get_up() == 1
    program.start()
    program.stop()
    assert program.program_status() == lp.PS.EXCEPTION and "CANCELLED" in program.exceptions[0]


def test_engine_refuses_cpu_tiles(unique_key):
    from numpywren_b200 import _capi
    _, _, _, program = build(unique_key)
    program.start()
    with pytest.raises(_capi.NpwError, match="no CPU execution path"):
        job_runner.lambdapack_run(program, timeout=5)
    assert program.program_status() == lp.PS.EXCEPTION


def test_lru_cache_and_busy_time():
    c = job_runner.LRUCache(max_items=2)
    c["a"] = 1; c["b"] = 2; _ = c["a"]; c["c"] = 3
    assert "a" in c and "c" in c and "b" not in c
    with pytest.raises(KeyError):
        c["b"]
    assert job_runner.calculate_busy_time([[0, 2], [1, 3], [5, 6]]) == [[0, 3], [5, 6]]


def test_fast_host_loop_keeps_the_reference_bookkeeping(unique_key, monkeypatch):
    """job_runner's loop uses post_op_node / node-id queue items (keys formatted once per node); the resulting state —
    node status, edge sums, counted edges, terminators, order of execution — must be exactly what the public
    post_op path (the reference's protocol, key for key) produces."""
    from numpywren_b200.alg_wrappers import cholesky

    def strip(d, h):
        return {k.replace(h, "H"): v for k, v in d.items()} if isinstance(d, dict) else {k.replace(h, "H") for k in d}

    # A: the engine loop with a recording engine (fast path)
    A1 = BigMatrix(unique_key("fa"), shape=(24, 24), shard_sizes=(4, 4), device="cpu")
    prog_a, _ = cholesky(A1)
    order_a = []
    monkeypatch.setattr(job_runner.TileEngine, "run_node", lambda self, node: order_a.append(node.key))
    monkeypatch.setattr(job_runner.TileEngine, "finish", lambda self: [])
    job_runner.prepare(prog_a)
    prog_a.start()
    out = job_runner.lambdapack_run(prog_a, timeout=60)
    assert prog_a.program_status() == lp.PS.SUCCESS and len(out["executed_messages"]) == len(prog_a.program.nodes)
    # B: the public pieces (dequeue -> post_op -> set_node_status), same priorities
    A2 = BigMatrix(unique_key("fb"), shape=(24, 24), shard_sizes=(4, 4), device="cpu")
    prog_b, _ = cholesky(A2)
    prio = job_runner.TileEngine(prog_b).priorities()
    compiled = prog_b.program
    prog_b._priority_fn = lambda e, v: prio[compiled.node(e, v).nid]
    prog_b.start()
    order_b = []
    while True:
        item = prog_b._dequeue()
        if item is None:
            break
        e, v = item
        assert prog_b.get_node_status(e, v) == lp.NS.READY
        prog_b.set_node_status(e, v, lp.NS.RUNNING)
        prog_b.post_op(e, v, lp.PS.SUCCESS, None)
        prog_b.set_node_status(e, v, lp.NS.FINISHED)
        order_b.append(compiled.node(e, v).key)
    assert order_a == order_b
    ha, hb = prog_a.hash, prog_b.hash
    assert strip(prog_a._node_status, ha) == strip(prog_b._node_status, hb)
    assert strip(prog_a._edge_sum, ha) == strip(prog_b._edge_sum, hb)
    assert strip(prog_a._edges_seen, ha) == strip(prog_b._edges_seen, hb)
    assert prog_a._terminators_done == prog_b._terminators_done
    assert prog_a._get("progress") == prog_b._get("progress") == len(compiled.nodes)
    # public accessors see the fast path's state
    n0 = prog_a.program.nodes[0]
    assert prog_a.get_node_status(n0.expr_idx, n0.var_values) == lp.NS.FINISHED


def test_i8emu_digit_cache_is_released_after_the_last_syrk(unique_key, monkeypatch):
    """The experimental int8 path caches a panel tile's digits for the syrks that read it (like invdiag); the reference
    count must reach zero exactly when the last of them has been enqueued — no leak, no early drop."""
    import torch
    from numpywren_b200 import kernels
    from numpywren_b200.alg_wrappers import cholesky
    from numpywren_b200.compiler import _tile_key

    class Ev:
        def record(self, stream): pass

    splits = []
    monkeypatch.setattr(torch.cuda, "Event", Ev)
    monkeypatch.setattr(kernels, "split_i8", lambda tile, digits: (splits.append(tile) or ("d%d" % len(splits), "e")))
    A = BigMatrix(unique_key("dc"), shape=(28, 28), shard_sizes=(4, 4), device="cpu")
    program, _ = cholesky(A)
    eng = job_runner.TileEngine(program)
    live_max = 0
    stream = object()
    for node in program.program.nodes:                       # program order is a valid enqueue order
        if node.call.compute_name != "syrk":
            continue
        keys = [_tile_key(m, idx) for m, idx in node.reads]
        d1 = eng._tile_digits(node, 1, ("tile", keys[1]), keys[1], stream)
        d2 = eng._tile_digits(node, 2, ("tile", keys[2]), keys[2], stream)
        assert d1 is not None and d2 is not None
        if keys[1] == keys[2]:
            assert d1 == d2                                   # diagonal update: one extraction serves both operands
        live_max = max(live_max, len(eng._digits))
        eng._release_digits(keys[1])
        eng._release_digits(keys[2])
    nb = 7
    assert len(eng._digits) == 0                              # everything released
    assert len(splits) == nb * (nb - 1) // 2                  # each panel tile O[j,i] (j > i) is split exactly once
    assert live_max <= nb * (nb - 1) // 2


@pytest.mark.parametrize("which", ["qr", "bdfac", "gemm", "gemm_kloop"])
def test_dead_tile_reclamation_never_drops_a_tile_that_is_still_needed(unique_key, which):
    """free_intermediates: walking the DAG in a valid order with the engine's own release logic, every written tile must
    still be in the store when a reader comes for it, inputs / outputs are never touched, and every intermediate that
    was read at all is gone at the end."""
    import torch
    from numpywren_b200 import alg_wrappers
    from numpywren_b200.compiler import _tile_key

    A = BigMatrix(unique_key("ra"), shape=(24, 24), shard_sizes=(4, 4), device="cpu")
    B = BigMatrix(unique_key("rb"), shape=(24, 24), shard_sizes=(4, 4), device="cpu")
    for m in (A, B):
        for bi in m.block_idxs:
            m._put_block_ref(torch.zeros(4, 4, dtype=torch.float64), *bi)
    program, meta = (getattr(alg_wrappers, which)(A, B) if which.startswith("gemm") else getattr(alg_wrappers, which)(A))
    for m in meta["outputs"] + meta["intermediates"]:
        m.free()
    cp = program.program
    eng = job_runner.TileEngine(program, free_intermediates=True)
    assert eng.free_intermediates
    keep = {id(m) for m in [A, B] + meta["outputs"]}
    starters = {(int(e), tuple(sorted((str(k), int(v)) for k, v in vv.items()))) for e, vv in cp.starters}
    ran = set()
    for node in cp.nodes:
        if node.key not in starters and not (node.parents and all(p in ran for p in node.parents)):
            continue
        refs, keys = [], []
        for (m, idx) in node.reads:
            ref = m._get_block_ref(*idx)
            if cp.writer_of(m, idx) is not None:
                assert ref is not None, f"{which}: tile {m.key}{list(idx)} was dropped before {node} read it"
            refs.append(ref)
            keys.append(_tile_key(m, idx))
        for (m, idx) in node.writes:
            m._put_block_ref(torch.zeros(1, 1, dtype=torch.float64), *idx)
        eng._release_dead_inputs(node, refs, keys, None)
        ran.add(node.nid)
    assert eng.freed_tiles > 0
    for m in [A, B] + meta["outputs"]:
        assert len(m._blocks_store) > 0                                        # never touched
    for m in meta["intermediates"]:
        if id(m) in keep:
            continue
        for idx in list(m._blocks_store):
            assert cp.num_readers(m, idx) == 0 or any(cp.nodes[r].nid not in ran for r in cp._readers[_tile_key(m, idx)]), \
                f"{which}: {m.key}{list(idx)} was read by every consumer but not reclaimed"
