"""The LambdaPACK compiler against Python itself: every program in algs.py is also *executed* as the plain Python
function it syntactically is, with recording stand-ins for the matrices and kernels; the trace of (kernel, reads,
writes) must equal the compiler's expanded node list, in order.  This pins the expression evaluator (ranges, ``**``,
``ceiling(log(.)/log(.))``, nested scopes) to Python/math semantics for a sweep of sizes."""
import inspect
import math
import textwrap

import pytest

from numpywren_b200 import algs, compiler
from numpywren_b200.matrix import BigMatrix

_n = [0]


def dummy(nd):
    _n[0] += 1
    return BigMatrix("sem_%d" % _n[0], shape=tuple([4096] * nd), shard_sizes=tuple([1] * nd), device="cpu")


class Rec:
    """Matrix stand-in: indexing returns a read token, item assignment records a write."""

    def __init__(self, name, trace):
        self.name, self.trace = name, trace

    def __getitem__(self, idx):
        idx = idx if isinstance(idx, tuple) else (idx,)
        return ("read", self.name, tuple(int(i) for i in idx))

    def __setitem__(self, idx, value):
        idx = idx if isinstance(idx, tuple) else (idx,)
        value["writes"].append((self.name, tuple(int(i) for i in idx)))


def native_trace(fn, arg_names, args):
    trace = []

    def kernel(name):
        def call(*a):
            node = {"fn": name, "reads": [(x[1], x[2]) for x in a if isinstance(x, tuple) and x and x[0] == "read"],
                    "writes": []}
            trace.append(node)
            return node
        return call

    class Multi(dict):
        pass

    ns = {"ceiling": lambda x: int(math.ceil(round(x, 9))), "floor": lambda x: int(math.floor(round(x, 9))),
          "log": math.log, "BigMatrix": BigMatrix}
    for k in ("chol", "trsm", "syrk", "gemm", "add_matrices", "identity", "qr_factor", "qr_factor_triangular", "qr_leaf",
              "qr_trailing_update", "lq_factor", "lq_leaf", "lq_trailing_update", "gemm_acc"):
        ns[k] = kernel(k)
    src = textwrap.dedent(inspect.getsource(fn))
    # tuple targets "A[..], B[..], C[..] = f(...)" assign the same call record to each target
    import ast

    class SplitTuple(ast.NodeTransformer):
        def visit_Assign(self, node):
            if isinstance(node.targets[0], ast.Tuple) and isinstance(node.value, ast.Call):
                tmp = ast.Name(id="__call", ctx=ast.Store())
                out = [ast.Assign(targets=[tmp], value=node.value)]
                for t in node.targets[0].elts:
                    out.append(ast.Assign(targets=[t], value=ast.Name(id="__call", ctx=ast.Load())))
                return out
            return node
    tree = ast.fix_missing_locations(SplitTuple().visit(ast.parse(src)))
    exec(compile(tree, "<dsl>", "exec"), ns)
    bound = [Rec(n, trace) if isinstance(a, BigMatrix) else a for n, a in zip(arg_names, args)]
    ns[fn.__name__](*bound)
    return trace


CASES = [
    ("CHOLESKY", lambda n: (dummy(2), dummy(2), dummy(3), n, 0), [1, 2, 3, 5, 9, 16]),
    ("CHOLESKY", lambda n: (dummy(2), dummy(2), dummy(3), n, 2), [4, 7]),
    ("GEMM", lambda n: (dummy(2), dummy(2), n, n + 1, max(1, n - 1), dummy(4), dummy(2)), [1, 2, 3, 4, 5, 6, 17]),
    ("GEMM_ACC", lambda n: (dummy(2), dummy(2), n, n + 1, max(1, n - 1), dummy(3), dummy(2)), [1, 2, 3, 5]),
    ("TSQR", lambda n: (dummy(2), dummy(2), dummy(2), dummy(2), n), [1, 2, 4, 8, 16, 32, 64]),
    ("QR", lambda n: (dummy(2), dummy(3), dummy(3), dummy(3), dummy(4), n, 0), [1, 2, 3, 4, 5, 8, 11]),
    ("BDFAC", lambda n: (dummy(2), dummy(3), dummy(3), dummy(4), dummy(3), dummy(3), dummy(3), dummy(4), dummy(3), n, 0),
     [2, 3, 4, 5, 8, 9]),
    ("BDFAC", lambda n: (dummy(2), dummy(3), dummy(3), dummy(4), dummy(3), dummy(3), dummy(3), dummy(4), dummy(3), n, 2), [4, 6]),
    ("SimpleTestLinear", lambda n: (dummy(2), dummy(2), n), [1, 3, 6]),
    ("SimpleTestLinear2", lambda n: (dummy(2), dummy(2), n), [2, 5]),
    ("SimpleTestNonLinear", lambda n: (dummy(3), dummy(1), n), [1, 2, 4, 8, 16]),
]


@pytest.mark.parametrize("prog,make,sizes", CASES)
def test_expansion_equals_native_python_execution(prog, make, sizes):
    fn = getattr(algs, prog)
    arg_names = list(inspect.signature(fn).parameters)
    for n in sizes:
        args = make(n)
        names = {id(a): nm for nm, a in zip(arg_names, args) if isinstance(a, BigMatrix)}
        p = compiler.lpcompile(fn)(*args)
        got = [{"fn": nd.call.compute_name, "reads": [(names[id(m)], idx) for m, idx in nd.reads],
                "writes": [(names[id(m)], idx) for m, idx in nd.writes]} for nd in p.nodes]
        want = native_trace(fn, arg_names, args)
        assert len(got) == len(want), (prog, n, len(got), len(want))
        for g, w in zip(got, want):
            assert g == w, (prog, n, g, w)


def test_non_power_of_two_trees_read_unwritten_tiles_like_the_reference():
    """TSQR with N = 3 reads Rs[0, 3], which nothing writes: the node has a missing parent (the reference would fail the
    S3 GET at run time); the compiler must still expand the program and report the dangling read as parentless."""
    A, V, T, R = dummy(2), dummy(2), dummy(2), dummy(2)
    p = compiler.lpcompile(algs.TSQR)(A, V, T, R, 3)
    merge = [n for n in p.nodes if n.expr_idx == 1 and n.var_values == {"level": 0, "j": 2}][0]
    assert merge.reads[1][1] == (0, 3) and p.writer_of(R, (0, 3)) is None
    assert len(merge.parents) == 1
