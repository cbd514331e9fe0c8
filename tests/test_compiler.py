"""LambdaPACK front end + DAG compiler against the reference's structural known answers
(tests/test_starters_terminators.py:14-43, tests/test_dependency_analyze.py:18-106) and against the
reference's own find_children / find_parents output captured in tests/golden/structure.json."""
import json
import os

import pytest

from numpywren_b200 import algs, compiler, exceptions, frontend
from numpywren_b200.matrix import BigMatrix

_n = [0]


def dummy_matrix(num_dims=2):
    _n[0] += 1
    return BigMatrix("dummy_%d" % _n[0], shape=tuple([1000] * num_dims), shard_sizes=tuple([1] * num_dims), device="cpu")


def key(node):
    return repr((int(node[0]), tuple(sorted((str(k), int(v)) for k, v in node[1].items()))))


@pytest.fixture(scope="module")
def structure(golden_dir):
    with open(os.path.join(golden_dir, "structure.json")) as f:
        return json.load(f)


def verify_program(program):
    """Reference tests/test_dependency_analyze.py:18-40: parent/child relations are mutually consistent."""
    for p_idx, loop_vars in compiler.walk_program(program):
        for c in compiler.find_children(program, p_idx, loop_vars):
            assert key((p_idx, loop_vars)) in [key(p) for p in compiler.find_parents(program, *c)]
        for p in compiler.find_parents(program, p_idx, loop_vars):
            assert key((p_idx, loop_vars)) in [key(c) for c in compiler.find_children(program, *p)]


def test_cholesky_starters_terminators_known_answer():
    # tests/test_starters_terminators.py:14-20
    program = compiler.lpcompile(algs.CHOLESKY)(dummy_matrix(), dummy_matrix(), dummy_matrix(3), 313, 0)
    assert compiler.find_starters(program, input_matrices=["I"]) == [(0, {})]
    assert len(compiler.find_terminators(program, output_matrices=["O"])) == 49141


def test_gemm_starters_terminators_known_answer():
    # tests/test_starters_terminators.py:33-43
    M = N = K = 4
    program = compiler.lpcompile(algs.GEMM)(dummy_matrix(), dummy_matrix(), M, N, K, dummy_matrix(4), dummy_matrix(3))
    assert len(compiler.find_starters(program, input_matrices=["A", "B"])) == M * N * K
    assert len(compiler.find_terminators(program, output_matrices=["Out"])) == M * N


@pytest.mark.parametrize("nb", [1, 2, 3, 4, 6, 8, 16])
def test_cholesky_counts_match_reference(structure, nb):
    ref = structure["structure"]["cholesky_%d" % nb]
    program = compiler.lpcompile(algs.CHOLESKY)(dummy_matrix(), dummy_matrix(), dummy_matrix(3), nb, 0)
    nodes = compiler.walk_program(program)
    assert len(nodes) == ref["nodes"] == nb + nb * (nb - 1) // 2 + (nb - 1) * nb * (nb + 1) // 6
    per = {}
    for e, _ in nodes:
        per[str(e)] = per.get(str(e), 0) + 1
    assert per == ref["per_expr"]
    assert [key(s) for s in compiler.find_starters(program, ["I"])] == ref["starters"]
    assert len(compiler.find_terminators(program, ["O"])) == ref["terminators"] == nb * (nb + 1) // 2


def test_gemm_and_tsqr_counts_match_reference(structure):
    for name, ref in structure["structure"].items():
        if name.startswith("gemm_"):
            M, N, K = (int(x) for x in name.split("_")[1:])
            p = compiler.lpcompile(algs.GEMM)(dummy_matrix(), dummy_matrix(), M, N, K, dummy_matrix(4), dummy_matrix(3))
            assert (len(p.nodes), len(compiler.find_starters(p, ["A", "B"])), len(compiler.find_terminators(p, ["Out"]))) == \
                (ref["nodes"], ref["starters"], ref["terminators"]), name
        elif name.startswith("tsqr_"):
            N = int(name.split("_")[1])
            p = compiler.lpcompile(algs.TSQR)(dummy_matrix(), dummy_matrix(), dummy_matrix(), dummy_matrix(), N)
            assert (len(p.nodes), len(compiler.find_starters(p, ["A"])), len(compiler.find_terminators(p, ["Rs"]))) == \
                (ref["nodes"], ref["starters"], ref["terminators"]), name


def test_qr_and_bdfac_counts_match_reference(structure):
    """Starters / terminators (and node counts) of algs.QR and algs.BDFAC as the reference's compiler reports them,
    including the size of its own known-answer test (tests/test_starters_terminators.py:22-31, M = 256)."""
    seen = 0
    for name, ref in structure["structure"].items():
        if name.startswith("qr_"):
            N = int(name.split("_")[1])
            p = compiler.lpcompile(algs.QR)(dummy_matrix(), dummy_matrix(3), dummy_matrix(3), dummy_matrix(3), dummy_matrix(4), N, 0)
            assert len(compiler.find_starters(p, ["I"])) == ref["starters"] == N, name
            assert len(compiler.find_terminators(p, ["Rs"])) == ref["terminators"], name
            if ref["nodes"] is not None:
                assert len(compiler.walk_program(p)) == ref["nodes"], name
            seen += 1
        elif name.startswith("bdfac_"):
            N = int(name.split("_")[1])
            p = compiler.lpcompile(algs.BDFAC)(dummy_matrix(), dummy_matrix(3), dummy_matrix(3), dummy_matrix(4), dummy_matrix(3),
                                               dummy_matrix(3), dummy_matrix(3), dummy_matrix(4), dummy_matrix(3), N, 0)
            assert (len(compiler.walk_program(p)), len(compiler.find_starters(p, ["I"])),
                    len(compiler.find_terminators(p, ["R_QR", "L_LQ"]))) == (ref["nodes"], ref["starters"], ref["terminators"]), name
            seen += 1
    assert seen >= 11


@pytest.mark.parametrize("which", ["cholesky_64_16", "gemm_64_16", "tsqr_256_32", "qr_28_7", "bdfac_16_4"])
def test_dag_edges_equal_reference_symbolic_analysis(structure, which):
    dag = structure[which]["dag"]
    if which.startswith("qr"):
        p = compiler.lpcompile_for_execution(algs.QR, inputs=["I"], outputs=["Rs"])(
            dummy_matrix(), dummy_matrix(3), dummy_matrix(3), dummy_matrix(3), dummy_matrix(4), 4, 0)
        assert (len(p.starters), p.num_terminators) == (structure[which]["num_starters"], structure[which]["num_terminators"])
    elif which.startswith("bdfac"):
        p = compiler.lpcompile_for_execution(algs.BDFAC, inputs=["I"], outputs=["R_QR", "L_LQ"])(
            dummy_matrix(), dummy_matrix(3), dummy_matrix(3), dummy_matrix(4), dummy_matrix(3), dummy_matrix(3), dummy_matrix(3),
            dummy_matrix(4), dummy_matrix(3), 4, 0)
        assert (len(p.starters), p.num_terminators) == (structure[which]["num_starters"], structure[which]["num_terminators"])
    elif which.startswith("cholesky"):
        p = compiler.lpcompile(algs.CHOLESKY)(dummy_matrix(), dummy_matrix(), dummy_matrix(3), 4, 0)
    elif which.startswith("gemm"):
        p = compiler.lpcompile(algs.GEMM)(dummy_matrix(), dummy_matrix(), 4, 4, 4, dummy_matrix(4), dummy_matrix())
    else:
        p = compiler.lpcompile(algs.TSQR)(dummy_matrix(), dummy_matrix(), dummy_matrix(), dummy_matrix(), 8)
    assert len(p.nodes) == len(dag)
    for n in p.nodes:
        ref = dag[key(n.ref)]
        assert sorted(key(c) for c in p.find_children(*n.ref)) == ref["children"]
        assert sorted(key(c) for c in p.find_parents(*n.ref)) == ref["parents"]


@pytest.mark.parametrize("prog,args", [
    ("SimpleTestLinear", (2, 2, 5)), ("SimpleTestLinear2", (2, 2, 5)), ("SimpleTestNonLinear", (3, 1, 8)),
    ("CHOLESKY", (2, 2, 3, 8, 0)), ("TSQR", (2, 2, 2, 2, 16)), ("GEMM", (2, 2, 4, 4, 4, 4, 2))])
def test_verify_program(prog, args):
    # tests/test_dependency_analyze.py:42-106
    bound = [dummy_matrix(a) if i < {"SimpleTestLinear": 2, "SimpleTestLinear2": 2, "SimpleTestNonLinear": 2, "CHOLESKY": 3,
                                      "TSQR": 4}.get(prog, 0) else a for i, a in enumerate(args)]
    if prog == "GEMM":
        bound = [dummy_matrix(2), dummy_matrix(2), 4, 4, 4, dummy_matrix(4), dummy_matrix(2)]
    verify_program(compiler.lpcompile(getattr(algs, prog))(*bound))


def test_cholesky_nb32_is_the_benchmark_dag():
    p = compiler.lpcompile_for_execution(algs.CHOLESKY, inputs=["I"], outputs=["O"])(
        dummy_matrix(), dummy_matrix(), dummy_matrix(3), 32, 0)
    per = {}
    for n in p.nodes:
        per[n.call.compute_name] = per.get(n.call.compute_name, 0) + 1
    assert per == {"chol": 32, "trsm": 496, "syrk": 5456}      # SURVEY §8 task counts
    assert p.starters == [(0, {})] and p.num_terminators == 528
    assert p.is_terminator(0) and p.is_terminator(4) and not p.is_terminator(5)
    # every S tile has exactly one reader: the in-place aliasing precondition
    for n in p.nodes:
        if n.call.compute_name == "syrk":
            m, idx = n.writes[0]
            assert p.num_readers(m, idx) == 1


def test_truncate_shrinks_the_program():
    full = compiler.lpcompile(algs.CHOLESKY)(dummy_matrix(), dummy_matrix(), dummy_matrix(3), 6, 0)
    trunc = compiler.lpcompile(algs.CHOLESKY)(dummy_matrix(), dummy_matrix(), dummy_matrix(3), 6, 2)
    ref4 = compiler.lpcompile(algs.CHOLESKY)(dummy_matrix(), dummy_matrix(), dummy_matrix(3), 4, 0)
    assert len(trunc.nodes) == len(ref4.nodes) < len(full.nodes)


def test_eval_expr_builds_read_call_write_block():
    from numpywren_b200 import kernels, lambdapack as lp
    O, I, S = dummy_matrix(), dummy_matrix(), dummy_matrix(3)
    p = compiler.lpcompile_for_execution(algs.CHOLESKY, ["I"], ["O"])(O, I, S, 4, 0)
    ib = p.eval_expr(5, {"i": 1, "j": 3, "k": 2})
    reads = [x for x in ib.instrs if isinstance(x, lp.RemoteRead)]
    calls = [x for x in ib.instrs if isinstance(x, lp.RemoteCall)]
    writes = [x for x in ib.instrs if isinstance(x, lp.RemoteWrite)]
    assert [(r.matrix, r.bidxs) for r in reads] == [(S, (1, 3, 2)), (O, (3, 1)), (O, (2, 1))]
    assert len(calls) == 1 and calls[0].compute is kernels.syrk
    assert [(w.matrix, w.bidxs) for w in writes] == [(S, (2, 3, 2))]
    assert [type(x) for x in ib.instrs] == [lp.RemoteRead] * 3 + [lp.RemoteCall, lp.RemoteWrite]


def test_exact_log_ceiling():
    for base in (2, 4):
        for k in range(1, 40):
            assert frontend._ceiling(frontend._log(base ** k) / frontend._log(base)) == k
            assert frontend._ceiling(frontend._log(base ** k + 1) / frontend._log(base)) == k + 1


def test_static_if_selects_branch():
    def prog(A: BigMatrix, B: BigMatrix, N: int):
        for i in range(N):
            if i % 2 == 0:
                B[i, 0] = identity(A[i, 0])
            else:
                B[i, 1] = identity(A[i, 1])
    p = compiler.lpcompile(prog)(dummy_matrix(), dummy_matrix(), 5)
    got = sorted((n.expr_idx, n.var_values["i"]) for n in p.nodes)
    assert got == [(0, 0), (0, 2), (0, 4), (1, 1), (1, 3)]


def test_float_args_kept_int_args_dropped():
    from numpywren_b200 import lambdapack as lp

    def prog(A: BigMatrix, B: BigMatrix, N: int):
        B[0, 0] = identity(A[0, 0], 2.5, N, 3)
    p = compiler.lpcompile(prog)(dummy_matrix(), dummy_matrix(), 7)
    ib = p.eval_expr(0, {})
    call = [x for x in ib.instrs if isinstance(x, lp.RemoteCall)][0]
    call.argv_instr[0].result = "tile"
    assert call._pyargs() == ["tile", 2.5]       # reference lambdapack.py:364-368


def test_rejections():
    def returns(A: BigMatrix, N: int):
        return A

    def while_loop(A: BigMatrix, N: int):
        for i in [1, 2]:
            A[i] = identity(A[i])

    def unknown_kernel(A: BigMatrix, N: int):
        A[0] = no_such_kernel(A[1])

    def redeclare(A: BigMatrix, N: int):
        x = 1
        x = 2
        A[0] = identity(A[x])

    def not_ssa(A: BigMatrix, B: BigMatrix, N: int):
        for i in range(N):
            B[0, 0] = identity(A[i, 0])

    with pytest.raises(exceptions.LambdaPackParsingException):
        compiler.lpcompile(returns)
    with pytest.raises(NotImplementedError):
        compiler.lpcompile(while_loop)
    with pytest.raises(Exception, match="unsupported function"):
        compiler.lpcompile(unknown_kernel)
    with pytest.raises(exceptions.LambdaPackParsingException):
        compiler.lpcompile(redeclare)
    with pytest.raises(Exception, match="SSA"):
        compiler.lpcompile(not_ssa)(dummy_matrix(), dummy_matrix(), 3).nodes
    with pytest.raises(exceptions.LambdaPackBackendGenerationException):
        compiler.lpcompile(algs.CHOLESKY)(dummy_matrix(), "not a matrix", dummy_matrix(3), 4, 0)
    with pytest.raises(AssertionError):
        compiler.lpcompile(algs.CHOLESKY)(dummy_matrix(), dummy_matrix())
