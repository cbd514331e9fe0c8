"""The REAL stream engine on the host (tests/_fakecuda.py: inert CUDA streams/events + the NumPy C-ABI double): every
program the package ships is run through alg_wrappers -> program.start() -> job_runner.lambdapack_run — native DAG
expansion, priorities, the fast host loop, in-place aliasing, invdiag hand-over, lower-only diagonal updates, multi-output
stores, default (parent_fn) tiles, dead-tile reclamation — and compared with the golden outputs of the unmodified
reference.  What this cannot see is the CUDA kernels themselves and real stream concurrency; the `-m gpu` tests do."""
import os

import numpy as np
import pytest
import torch

import _fakecuda
from numpywren_b200 import alg_wrappers, binops, job_runner, qr
from numpywren_b200 import lambdapack as lp
from numpywren_b200.matrix import BigMatrix
from numpywren_b200.matrix_init import shard_matrix
from oracle import npw_oracle as orc


@pytest.fixture
def engine(monkeypatch):
    prev = qr.get_qr_semantics()
    lib = _fakecuda.install(monkeypatch)
    yield lib
    qr.set_qr_semantics(prev)


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def run(program, expect=lp.PS.SUCCESS, **kw):
    program.start()
    out = job_runner.lambdapack_run(program, timeout=120, **kw)
    assert program.program_status() == expect, program.exceptions
    return out


def cpu_matrix(key, X, b, **kw):
    A = BigMatrix(key, shape=X.shape, shard_sizes=(b, b), device="cpu", **kw)
    A.free()
    shard_matrix(A, X)
    return A


@pytest.mark.parametrize("name", ["cholesky_64_8", "cholesky_64_16", "cholesky_60_16", "cholesky_64_32_lam"])
@pytest.mark.parametrize("inplace", [True, False])
def test_cholesky_golden_through_the_engine(engine, golden_dir, unique_key, name, inplace):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    n, b, lam = int(g["n"]), int(g["b"]), float(g["lambdav"])
    A = cpu_matrix(unique_key(name), g["A"], b, lambdav=lam)
    program, meta = alg_wrappers.cholesky(A)
    for m in meta["outputs"] + meta["intermediates"]:
        m.free()
    res = run(program, inplace=inplace)
    assert len(res["executed_messages"]) == int(g["nnodes"]) and program.program.expanded_by == "native"
    L = meta["outputs"][0].numpy()
    assert rel(L, g["L"]) < 1e-12
    assert np.array_equal(A.numpy(), g["A"] + lam * np.eye(n))              # inputs are never consumed by default
    S = meta["intermediates"][0]
    if inplace:
        assert len(S.block_idxs_exist) == 0                                  # every S version was aliased in place
    else:
        for k in g.files:
            if k.startswith("S_"):
                i, j, kk = (int(x) for x in k.split("_")[1:])
                got, ref = S.get_block(i, j, kk).numpy(), g[k].reshape(S.get_block(i, j, kk).shape)
                if j == kk:
                    got, ref = np.tril(got), np.tril(ref)
                assert rel(got, ref) < 1e-12, k
    # chol handed its by-product to the trsms of its column, and the diagonal updates used the lower-only entry point
    kinds = [c[0] for c in engine.calls]
    assert "potrf" in kinds and ("trsm_rlt" in kinds or n == b)


def test_large_tiles_use_lower_only_diagonal_updates_and_consume_inputs(engine, unique_key):
    n, b = 768, 256
    rs = np.random.RandomState(0)
    x = rs.randn(n, 64)
    a = x @ x.T + n * np.eye(n)
    A = cpu_matrix(unique_key("big"), a, b)
    program, meta = alg_wrappers.cholesky(A)
    for m in meta["outputs"] + meta["intermediates"]:
        m.free()
    run(program, consume_inputs=True)
    L = meta["outputs"][0].numpy()
    assert rel(L, np.linalg.cholesky(a)) < 1e-12
    assert any(c[0] == "syrk_lower" for c in engine.calls)
    assert len(A.block_idxs_exist) < len(A.block_idxs)                       # input buffers were re-used


def test_not_positive_definite_is_reported_after_the_drain(engine, unique_key):
    bad = np.eye(64)
    bad[50, 50] = -1.0
    A = cpu_matrix(unique_key("bad"), bad, 16)
    program, meta = alg_wrappers.cholesky(A)
    program.start()
    with pytest.raises(np.linalg.LinAlgError):
        job_runner.lambdapack_run(program, timeout=60)
    assert program.program_status() == lp.PS.EXCEPTION


@pytest.mark.parametrize("name", ["gemm_64_16", "gemm_32_16"])
@pytest.mark.parametrize("free", [False, True])
def test_gemm_golden_through_the_engine(engine, golden_dir, unique_key, name, free):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    b = int(g["b"])
    A = cpu_matrix(unique_key("gA"), g["A"], b)
    B = cpu_matrix(unique_key("gB"), g["B"], b)
    program, meta = alg_wrappers.gemm(A, B)
    for m in meta["outputs"] + meta["intermediates"]:
        m.free()
    res = run(program, free_intermediates=free)
    assert len(res["executed_messages"]) == int(g["nnodes"])
    assert rel(meta["outputs"][0].numpy(), g["C"]) < 1e-13
    if free:
        assert program._engine.freed_tiles > 0 and len(meta["intermediates"][0]._blocks_store) == 0


@pytest.mark.parametrize("name", ["tsqr_256_32", "tsqr_128_16"])
def test_tsqr_golden_through_the_engine(engine, golden_dir, unique_key, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    m_, b, nlev = int(g["m"]), int(g["b"]), int(g["nlev"])
    X = BigMatrix(unique_key("tX"), shape=(m_, b), shard_sizes=(b, b), device="cpu")
    X.free()
    shard_matrix(X, g["X"])
    program, meta = alg_wrappers.tsqr(X)
    for mm in meta["outputs"]:
        mm.free()
    res = run(program)
    assert len(res["executed_messages"]) == int(g["nnodes"])
    assert rel(meta["outputs"][0].get_block(nlev, 0).numpy(), g["R"]) < 1e-12


def _check(g, mats, numeric=lambda name, idx: True):
    n = 0
    for k in g.files:
        name = next((m for m in mats if k.startswith(m + "_")), None)
        if name is None:
            continue
        idx = tuple(int(x) for x in k[len(name) + 1:].split("_"))
        got = mats[name]._blocks_store[idx].numpy()
        assert got.size == g[k].size
        if numeric(name, idx):
            assert np.abs(got.reshape(g[k].shape) - g[k]).max() <= 1e-10 * max(1.0, np.abs(g[k]).max()), k
        n += 1
    assert n == sum(len(m._blocks_store) for m in mats.values())


@pytest.mark.parametrize("name", ["qr_28_7", "qr_16_8", "qr_24_8"])
def test_qr_golden_through_the_engine(engine, golden_dir, unique_key, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    qr.set_qr_semantics("reference")
    A = cpu_matrix(unique_key("qA"), g["X"], int(g["b"]))
    program, meta = alg_wrappers.qr(A)
    mats = dict(zip(["Rs", "Vs", "Ts", "Ss"], meta["outputs"] + meta["intermediates"]))
    for m in mats.values():
        m.free()
    res = run(program)
    assert len(res["executed_messages"]) == int(g["nnodes"])
    _check(g, mats)


@pytest.mark.parametrize("name", ["bdfac_16_4", "bdfac_16_4_trunc2", "bdfac_15_5"])
def test_bdfac_golden_through_the_engine(engine, golden_dir, unique_key, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    qr.set_qr_semantics("reference")
    A = cpu_matrix(unique_key("bA"), g["X"], int(g["b"]))
    program, meta = alg_wrappers.bdfac(A, truncate=int(g["truncate"]))
    mats = dict(zip(["L_LQ", "R_QR", "S_LQ", "S_QR", "T_QR", "V_QR", "V_LQ", "T_LQ"], meta["outputs"] + meta["intermediates"]))
    for m in mats.values():
        m.free()
    # a truncated BDFAC never finishes — in the reference either: its last statement writes an output (so it counts as
    # a terminator, 14 in all) but reads a tile the truncated loops never produce, so it never becomes ready and the
    # terminator count stops at 13.  The runner returns when nothing is runnable; the status stays RUNNING.
    res = run(program, expect=lp.PS.RUNNING if int(g["truncate"]) else lp.PS.SUCCESS)
    assert len(res["executed_messages"]) == int(g["nnodes"]) and program.queue_depth() == 0
    _check(g, mats)


def test_qr_householder_with_reclamation_through_the_engine(engine, unique_key):
    qr.set_qr_semantics("householder")
    n, b = 96, 16
    X = np.random.RandomState(4).randn(n, n)
    A = cpu_matrix(unique_key("qh"), X, b)
    program, meta = alg_wrappers.qr(A)
    for m in meta["outputs"] + meta["intermediates"]:
        m.free()
    run(program, free_intermediates=True)
    Rs = meta["outputs"][0]
    nb = n // b
    R = np.zeros((n, n))
    for i in range(nb):
        for k in range(i, nb):
            R[i * b:(i + 1) * b, k * b:(k + 1) * b] = Rs.get_block(i, k, 0).numpy()
    assert np.abs(np.abs(R) - np.abs(np.linalg.qr(X)[1])).max() < 1e-10
    assert program._engine.freed_tiles > 0


def test_gemm_kloop_accumulates_in_place_through_the_engine(engine, unique_key):
    rs = np.random.RandomState(6)
    a, b = rs.randn(300, 260), rs.randn(260, 200)
    A = cpu_matrix(unique_key("kA"), a, 128)
    B = cpu_matrix(unique_key("kB"), b, 128)
    program, meta = alg_wrappers.gemm_kloop(A, B)
    for m in meta["outputs"] + meta["intermediates"]:
        m.free()
    run(program)
    Acc, Out = meta["intermediates"][0], meta["outputs"][0]
    assert rel(Out.numpy(), a @ b) < 1e-13
    assert len(Acc._blocks_store) == len(Out._blocks_store) == 6        # only the last version of every output tile is left
    assert np.array_equal(A.numpy(), a) and np.array_equal(B.numpy(), b)


@pytest.mark.parametrize("digits,tol", [(5, 1e-9), (6, 1e-11), (8, 1e-14)])
def test_experimental_i8emu_mode_through_the_engine(engine, unique_key, monkeypatch, digits, tol):
    """NPW_B200_SYRK=i8emu: the engine extracts a panel tile's int8 digits once, reuses them for every syrk of its block
    row / column, releases them after the last one, and the factor stays within the accuracy the prototype predicts.
    (The C-ABI double restates npw_split_i8_f64 / npw_syrk_i8emu_f64 from the header; the tcgen05 kernel itself has
    never run.)"""
    monkeypatch.setenv("NPW_B200_SYRK", "i8emu")
    monkeypatch.setenv("NPW_B200_I8_DIGITS", str(digits))
    n, b = 640, 128
    nb = n // b
    a = np.block([[orc.spd_tile(j, k, b, n, width=64) for k in range(nb)] for j in range(nb)])
    A = cpu_matrix(unique_key("i8"), a, b)
    program, meta = alg_wrappers.cholesky(A)
    for m in meta["outputs"] + meta["intermediates"]:
        m.free()
    run(program)
    L = meta["outputs"][0].numpy()
    assert rel(L, np.linalg.cholesky(a)) < tol
    kinds = [c[0] for c in engine.calls]
    assert kinds.count("split_i8") == nb * (nb - 1) // 2          # once per panel tile O[j,i], j > i
    assert kinds.count("syrk_i8emu") == (nb - 1) * nb * (nb + 1) // 6 and "syrk" not in kinds and "syrk_lower" not in kinds
    assert any(c[0] == "syrk_i8emu" and c[5] == 1 for c in engine.calls)   # diagonal updates: lower-only
    assert len(program._engine._digits) == 0                      # every digit cache entry was released


def _host_bench_ctx(bench):
    """bench.Ctx without a GPU / process group: one rank, tiles on the host device."""
    ctx = object.__new__(bench.Ctx)
    ctx.world, ctx.rank, ctx.local_rank, ctx.device, ctx.grid = 1, 0, 0, torch.device("cpu"), None
    return ctx


def test_bench_gpu_step_flow_on_the_host(engine, monkeypatch):
    """bench.py's timed step (resident input -> cholesky() -> prepare -> start -> lambdapack_run), its residual check and
    its whole-program parity check against the oracle, with the host harness: the flow the driver runs at round end,
    minus the device."""
    import bench
    monkeypatch.setattr(bench.torch.cuda, "synchronize", lambda device=None: None)
    ctx = _host_bench_ctx(bench)
    wl = bench.CholeskyWorkload(ctx, 512, 128)
    ms, launches, plan_s, A, program, meta = bench.cholesky_step(ctx, wl, streams=4)
    assert program.program_status() == lp.PS.SUCCESS and launches > 0
    assert getattr(program, "_engine", None) is not None and program._engine.n_streams == 4
    assert bench.residual_check(ctx, wl, meta["outputs"][0]) < 1e-13
    bench.free_all(A, *meta["outputs"], *meta["intermediates"])
    t, L, tiles, cores = bench.cpu_cholesky_sample(512, 128)
    par = bench.parity_vs_oracle(ctx, 512, 128, 4, L, tiles)
    assert par["ok"] and par["rel_fro"] < 1e-13 and par["tiles_compared"] == 10
    L[(3, 1)] = L[(3, 1)] + 1e-6                      # the check must be able to fail
    assert not bench.parity_vs_oracle(ctx, 512, 128, 4, L, tiles)["ok"]
    tr = bench.utilisation_trace(ctx, wl, 4, slices=4)
    assert len(tr["busy_fraction_per_rank"]) == 1 and set(tr["rank0_kernel_ms"]) == {"chol", "trsm", "syrk"}


def test_bench_reference_arm_composes_the_same_workload(capsys, monkeypatch):
    """--impl reference: same config/metric as the GPU arm, step time composed from the oracle's kernel times by the
    task counts of the workload, all host cores even under torchrun's OMP_NUM_THREADS=1."""
    import json
    import bench
    monkeypatch.setenv("OMP_NUM_THREADS", "1")
    args = type("A", (), dict(workload="cholesky", n=1024, tile=128, gpus=2, steps=2, warmup=1, cpu_n=512, cols=512))()
    assert bench.run_reference_arm(args) == 0
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["config"]["workload"] == bench.workload_name(args)
    ks = line["config"]["kernel_seconds"]
    c, r, s = bench.chol_task_counts(8)
    assert (c, r, s) == (8, 28, 84)
    assert abs(line["ms_per_step"] - 1e3 * (c * ks["chol"] + r * ks["trsm"] + s * ks["syrk"])) < 1e-6 * line["ms_per_step"]
    assert line["cpu_baseline"]["cores"] == (os.cpu_count() or 1) and line["e2e"]["value"] == line["value"]
    assert line["config"]["cross_check"]["measured_end_to_end_s"] > 0


def test_duplicate_and_premature_messages_are_harmless(engine, golden_dir, unique_key):
    """Reference tests/test_failures.py:24-120 re-delivers random task messages and expects the same factor: a finished
    node is skipped (repeated_finish), a node whose parents have not all finished is not run early (not_ready) — which
    matters doubly here because finished inputs may already have been overwritten in place."""
    import random
    g = np.load(os.path.join(golden_dir, "cholesky_64_8.npz"))
    A = cpu_matrix(unique_key("dup"), g["A"], 8)
    program, meta = alg_wrappers.cholesky(A)
    for m in meta["outputs"] + meta["intermediates"]:
        m.free()
    job_runner.prepare(program)
    nodes = program.program.nodes
    rnd = random.Random(0)
    program.start()
    orig = job_runner.TileEngine.run_node
    injected = [0]

    def run_and_inject(self, node):
        orig(self, node)
        for _ in range(2):                                   # after every task, re-deliver two random messages
            victim = nodes[rnd.randrange(len(nodes))]
            program._enqueue(victim.expr_idx, victim.var_values, priority=rnd.randrange(100))
            injected[0] += 1
    job_runner.TileEngine.run_node = run_and_inject
    try:
        res = job_runner.lambdapack_run(program, timeout=120)
    finally:
        job_runner.TileEngine.run_node = orig
    assert program.program_status() == lp.PS.SUCCESS
    assert len(res["executed_messages"]) == len(nodes)                       # every node ran exactly once
    assert program._get("repeated_finish") + program._get("not_ready") == injected[0] - program.queue_depth() > 0
    assert rel(meta["outputs"][0].numpy(), g["L"]) < 1e-12


def test_several_runner_threads_on_one_program(engine, golden_dir, unique_key):
    """Reference tests/test_alg_correctness.py:160-187 / test_job_runner.py:64-91 submit several lambdapack_run workers for
    one program.  Here workers are threads of one process sharing the engine: together they run every node exactly once."""
    import concurrent.futures as fs
    g = np.load(os.path.join(golden_dir, "cholesky_64_8.npz"))
    A = cpu_matrix(unique_key("thr"), g["A"], 8)
    program, meta = alg_wrappers.cholesky(A)
    for m in meta["outputs"] + meta["intermediates"]:
        m.free()
    job_runner.prepare(program)
    program.start()
    with fs.ThreadPoolExecutor(3) as ex:
        futs = [ex.submit(job_runner.lambdapack_run, program, timeout=120) for _ in range(3)]
        outs = [f.result() for f in futs]
    program.wait()
    assert program.program_status() == lp.PS.SUCCESS
    assert sum(len(o["executed_messages"]) for o in outs) == len(program.program.nodes)
    assert rel(meta["outputs"][0].numpy(), g["L"]) < 1e-12
    assert program.get_up() == 0


@pytest.mark.parametrize("threads", [1, 3])
def test_eager_program_runs_the_child_post_op_hands_back(engine, golden_dir, unique_key, threads):
    """LambdaPackProgram(..., eager=True) (reference lambdapack.py:587-592): post_op pops one ready child and returns it
    as the next operator instead of queueing it; the reference runner executes it next (job_runner.py:113-139).  The
    engine loop must do the same, otherwise that child is READY for ever and the program stalls in RUNNING."""
    import concurrent.futures as fs
    g = np.load(os.path.join(golden_dir, "cholesky_64_8.npz"))
    A = cpu_matrix(unique_key("eager"), g["A"], 8)
    program, meta = alg_wrappers.cholesky(A)
    program.eager = True
    for m in meta["outputs"] + meta["intermediates"]:
        m.free()
    job_runner.prepare(program)
    program.start()
    with fs.ThreadPoolExecutor(threads) as ex:
        outs = [f.result() for f in [ex.submit(job_runner.lambdapack_run, program, timeout=120) for _ in range(threads)]]
    assert program.program_status() == lp.PS.SUCCESS
    assert sum(len(o["executed_messages"]) for o in outs) == len(program.program.nodes)
    assert rel(meta["outputs"][0].numpy(), g["L"]) < 1e-12
    assert program.get_up() == 0 and program._runner_active == 0


def test_success_is_published_by_the_last_runner_only(engine, golden_dir, unique_key):
    """A runner that drained the device early must not publish SUCCESS for a peer whose terminator is still being issued:
    the status flips exactly when the last runner leaves, after that runner's own finish()."""
    g = np.load(os.path.join(golden_dir, "cholesky_64_8.npz"))
    A = cpu_matrix(unique_key("last"), g["A"], 8)
    program, meta = alg_wrappers.cholesky(A)
    for m in meta["outputs"] + meta["intermediates"]:
        m.free()
    job_runner.prepare(program)
    program.start()
    with program._lock:
        program._runner_active += 1          # a second runner is "still issuing"
    job_runner.lambdapack_run(program, timeout=120)
    assert program._all_terminators_done and program.program_status() == lp.PS.RUNNING
    with program._lock:
        program._runner_active -= 1
    job_runner.lambdapack_run(program, timeout=120)          # the last runner: nothing left to issue, drains, publishes
    assert program.program_status() == lp.PS.SUCCESS
