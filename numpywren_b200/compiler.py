"""LambdaPACK compiler: IR → fully expanded tile DAG.

Public surface mirrors reference numpywren/compiler.py: ``lpcompile`` (:25-38),
``lpcompile_for_execution`` (:40-47), ``CompiledLambdaPackProgram`` with
``starters / num_terminators / find_children / find_parents / is_terminator / eval_expr``
(:49-69), and the module functions ``find_starters / find_terminators / find_parents /
find_children / walk_program`` (:595-650, 709-732, 780-791) used by the reference's tests.

The reference answers find_children / find_parents by *symbolic* index matching with sympy
at run time (6-11 ms per node) because its workers are stateless.  Here the
whole program runs inside one host process next to the GPUs, so the program is expanded
ONCE: every loop nest is enumerated on concrete values, every node records the tiles it
reads and writes, and edges come from a hash join  written-tile → readers.  For SSA
programs (each tile written once — the LambdaPACK contract, reference compiler.py:617-619)
this is exactly the relation the symbolic solver computes.  Node ids are the reference's:
``(expr_idx, {loop_var: value})`` with expr_idx numbering remote calls in source order.
"""
from __future__ import annotations

import time
from typing import Any, Dict, List, Optional, Sequence, Tuple

from . import exceptions, frontend
from .frontend import Assign, For, If, IndexExpr, RemoteCallAbstract

Node = Tuple[int, Dict[str, int]]


def _is_bigmatrix(x) -> bool:
    return hasattr(x, "get_block") and hasattr(x, "shard_sizes")


def _freeze(var_values: Dict[str, int]) -> Tuple[Tuple[str, int], ...]:
    return tuple(sorted((str(k), int(v)) for k, v in var_values.items()))


class ExpandedNode:
    """One tile task: a remote call with all loop variables bound."""
    __slots__ = ("nid", "expr_idx", "var_values", "call", "reads", "scalars", "arg_layout", "writes", "children",
                 "parents", "key")

    def __init__(self, nid, expr_idx, var_values, call, key=None):
        self.nid = nid
        self.expr_idx = expr_idx
        self.var_values = var_values
        self.call = call
        self.reads: List[Tuple[Any, Tuple[int, ...]]] = []     # (matrix, block idx) in argument order
        self.scalars: List[Any] = []
        self.arg_layout: List[Tuple[str, int]] = []            # ("read", i) | ("scalar", i) in DSL order
        self.writes: List[Tuple[Any, Tuple[int, ...]]] = []
        self.children: List[int] = []
        self.parents: List[int] = []
        self.key = key if key is not None else (expr_idx, _freeze(var_values))

    @property
    def ref(self) -> Node:
        return (self.expr_idx, dict(self.var_values))

    def __repr__(self):
        return f"Node({self.expr_idx}, {self.var_values}, {self.call.compute_name})"


def _tile_key(matrix, idx):
    """Identity of a stored tile: views of one matrix alias (reference cache key, lambdapack.py:244)."""
    true_idx = matrix.true_block_idx(*idx) if hasattr(matrix, "true_block_idx") else tuple(idx)
    return (getattr(matrix, "bucket", None), getattr(matrix, "key", id(matrix)), tuple(int(i) for i in true_idx))


class CompiledLambdaPackProgram:
    """A LambdaPACK program bound to concrete matrices/ints, expanded into a DAG on first use."""

    def __init__(self, fdef: frontend.FuncDef, args: Sequence[Any], inputs: Sequence[str], outputs: Sequence[str]):
        if len(fdef.args) != len(args):
            raise AssertionError("function {0} expected {1} args got {2}".format(fdef.name, len(fdef.args), len(args)))
        self.fdef = fdef
        self.inputs = list(inputs) if inputs is not None else []
        self.outputs = list(outputs) if outputs is not None else []
        self.scope: Dict[str, Any] = {}
        for name, typ, val in zip(fdef.args, fdef.arg_types, args):
            if typ == "BigMatrix" and not _is_bigmatrix(val):
                raise exceptions.LambdaPackBackendGenerationException(
                    "arg {0} wrong type expected BigMatrix got {1}".format(name, type(val).__name__))
            if typ == "int" and (isinstance(val, bool) or not isinstance(val, (int,)) and not hasattr(val, "__index__")):
                raise exceptions.LambdaPackBackendGenerationException(
                    "arg {0} wrong type expected int got {1}".format(name, type(val).__name__))
            if typ == "float" and not isinstance(val, (int, float)):
                raise exceptions.LambdaPackBackendGenerationException(
                    "arg {0} wrong type expected float got {1}".format(name, type(val).__name__))
            self.scope[name] = int(val) if typ == "int" else val
        # abstract remote calls in source order (reference BackendGenerate.remote_calls, frontend.py:779-784)
        self.remote_calls: Dict[int, RemoteCallAbstract] = {}
        self._index_calls(fdef.body)
        self._nodes: Optional[List[ExpandedNode]] = None
        self._by_key: Dict[Any, int] = {}
        self.expand_time = 0.0
        self._starters: Optional[List[Node]] = None
        self._num_terminators: Optional[int] = None

    # ------------------------------------------------------------------ structure
    def _index_calls(self, body):
        for s in body:
            if isinstance(s, RemoteCallAbstract):
                self.remote_calls[len(self.remote_calls)] = s
            elif isinstance(s, For):
                self._index_calls(s.body)
            elif isinstance(s, If):
                self._index_calls(s.body)
                self._index_calls(s.elseBody)

    def _matrix(self, name):
        m = self.scope.get(name)
        if m is None or not _is_bigmatrix(m):
            raise exceptions.LambdaPackBackendGenerationException(f"{name} is not a BigMatrix argument of {self.fdef.name}")
        return m

    def _contains(self, stmt) -> set:
        """Call indices syntactically inside a statement (memoised on the IR node)."""
        cached = self._contain_cache.get(id(stmt))
        if cached is not None:
            return cached
        out = set()
        if isinstance(stmt, RemoteCallAbstract):
            out.add(self._call_idx[id(stmt)])
        elif isinstance(stmt, For):
            for b in stmt.body:
                out |= self._contains(b)
        elif isinstance(stmt, If):
            for b in list(stmt.body) + list(stmt.elseBody):
                out |= self._contains(b)
        self._contain_cache[id(stmt)] = out
        return out

    def enumerate_instances(self, selected) -> List[Node]:
        """Instances ``(expr_idx, {loop vars})`` of the selected calls only, in program order, WITHOUT
        expanding the rest of the program (loops that contain no selected call are skipped).  This is what
        starters / terminators need: e.g. CHOLESKY(313) has 5.2 M nodes but only 49 141 terminators."""
        selected = set(selected)
        if not hasattr(self, "_call_idx"):
            self._call_idx = {id(c): i for i, c in self.remote_calls.items()}
            self._contain_cache = {}
        ints = {k: v for k, v in self.scope.items() if not _is_bigmatrix(v)}
        out: List[Node] = []

        def walk(body, env, loop_vars):
            for st in body:
                if isinstance(st, Assign):
                    env[st.name] = st.rhs.eval(env)
                    continue
                if not (self._contains(st) & selected):
                    continue
                if isinstance(st, RemoteCallAbstract):
                    out.append((self._call_idx[id(st)], dict(loop_vars)))
                elif isinstance(st, For):
                    lo, hi, stp = int(st.min.eval(env)), int(st.max.eval(env)), int(st.step.eval(env))
                    for v in range(lo, hi, stp):
                        env2 = dict(env)
                        env2[st.var] = v
                        lv = dict(loop_vars)
                        lv[st.var] = v
                        walk(st.body, env2, lv)
                elif isinstance(st, If):
                    walk(st.body if st.cond.eval(env) else st.elseBody, dict(env), loop_vars)

        walk(self.fdef.body, dict(ints), {})
        return out

    def _expand_native(self) -> bool:
        """Expansion by libnpw_dag (csrc/npw_dag.cpp).  Returns False when the Python expander below must run: library
        not built, program outside the native subset, or an error for which Python raises its own exception."""
        from . import _dag_native
        t0 = time.time()
        arr = _dag_native.expand(self)
        if arr is None:
            return False
        names, mats = arr["slot_names"], arr["matrices"]
        nt = arr["n_tiles"]
        tio, tix, tm = arr["tile_idx_off"].tolist(), arr["tile_idx"].tolist(), arr["tile_matrix"].tolist()
        from .matrix import BigMatrix
        tiles = [(mats[tm[t]], tuple(tix[tio[t]:tio[t + 1]])) for t in range(nt)]
        # _tile_key of a plain BigMatrix is (bucket, key, idx): skip the per-tile method calls for those
        plain = [type(m) is BigMatrix for m in mats]
        head = [(getattr(m, "bucket", None), getattr(m, "key", id(m))) for m in mats]
        keys = [head[tm[t]] + (tiles[t][1],) if plain[tm[t]] else _tile_key(*tiles[t]) for t in range(nt)]
        expr = arr["node_expr"].tolist()
        vo, vs, vv = arr["var_off"].tolist(), arr["var_slot"].tolist(), arr["var_val"].tolist()
        ro, rt = arr["read_off"].tolist(), arr["read_tile"].tolist()
        wo, wt = arr["write_off"].tolist(), arr["write_tile"].tolist()
        co, cn = arr["child_off"].tolist(), arr["child"].tolist()
        po, pn = arr["parent_off"].tolist(), arr["parent"].tolist()
        nodes: List[ExpandedNode] = []
        readers: Dict[Any, List[int]] = {}
        order_of: Dict[int, List[int]] = {}      # per remote call: positions of its loop variables sorted by name
        for nid in range(arr["n_nodes"]):
            e = expr[nid]
            call = self.remote_calls[e]
            lo, hi = vo[nid], vo[nid + 1]
            ent = order_of.get(e)
            if ent is None:
                vnames = [names[vs[a]] for a in range(lo, hi)]
                order = sorted(range(hi - lo), key=lambda a: vnames[a])
                ent = order_of[e] = (vnames, order, [vnames[a] for a in order])
            vnames, order, snames = ent
            vals = vv[lo:hi]
            key = (e, tuple(zip(snames, [vals[a] for a in order])))
            node = ExpandedNode(nid, e, dict(zip(vnames, vals)), call, key)
            node.reads = [tiles[t] for t in rt[ro[nid]:ro[nid + 1]]]
            node.arg_layout = [("read", a) for a in range(len(node.reads))]
            node.writes = [tiles[t] for t in wt[wo[nid]:wo[nid + 1]]]
            node.children = cn[co[nid]:co[nid + 1]]
            node.parents = pn[po[nid]:po[nid + 1]]
            for t in rt[ro[nid]:ro[nid + 1]]:
                readers.setdefault(keys[t], []).append(nid)
            nodes.append(node)
        tw = arr["tile_writer"].tolist()
        self._nodes = nodes
        self._by_key = {n.key: n.nid for n in nodes}
        self._writer = {keys[t]: tw[t] for t in range(nt) if tw[t] >= 0}
        self._readers = readers
        self.expand_time = time.time() - t0
        self.expanded_by = "native"
        return True

    def _expand(self):
        if self._expand_native():
            return
        self.expanded_by = "python"
        t0 = time.time()
        nodes: List[ExpandedNode] = []
        call_idx = {id(c): i for i, c in self.remote_calls.items()}
        ints = {k: v for k, v in self.scope.items() if not _is_bigmatrix(v)}

        def eval_index(ie: IndexExpr, env):
            m = self._matrix(ie.matrix_name)
            idx = []
            for e in ie.indices:
                v = e.eval(env)
                if isinstance(v, float):
                    if not v.is_integer():
                        raise exceptions.LambdaPackBackendGenerationException(
                            f"non-integer block index {v} in {ie.matrix_name}[{e.src}]")
                    v = int(v)
                idx.append(int(v))
            return m, tuple(idx)

        def walk(body, env, loop_vars):
            for s in body:
                if isinstance(s, RemoteCallAbstract):
                    node = ExpandedNode(len(nodes), call_idx[id(s)], dict(loop_vars), s)
                    for a in s.args:
                        if isinstance(a, IndexExpr):
                            node.arg_layout.append(("read", len(node.reads)))
                            node.reads.append(eval_index(a, env))
                        else:
                            node.arg_layout.append(("scalar", len(node.scalars)))
                            node.scalars.append(a.eval(env))
                    for o in s.output:
                        node.writes.append(eval_index(o, env))
                    nodes.append(node)
                elif isinstance(s, Assign):
                    env[s.name] = s.rhs.eval(env)
                elif isinstance(s, For):
                    lo, hi, st = s.min.eval(env), s.max.eval(env), s.step.eval(env)
                    lo, hi, st = int(lo), int(hi), int(st)
                    if st == 0:
                        raise exceptions.LambdaPackBackendGenerationException("range() step must not be zero")
                    for v in range(lo, hi, st):
                        env2 = dict(env)
                        env2[s.var] = v
                        lv = dict(loop_vars)
                        lv[s.var] = v
                        walk(s.body, env2, lv)
                elif isinstance(s, If):
                    walk(s.body if s.cond.eval(env) else s.elseBody, dict(env), loop_vars)

        walk(self.fdef.body, dict(ints), {})

        writer: Dict[Any, int] = {}
        readers: Dict[Any, List[int]] = {}
        for n in nodes:
            for (m, idx) in n.writes:
                k = _tile_key(m, idx)
                if k in writer:
                    raise Exception("Invalid Program Graph, LambdaPackPrograms must be SSA "
                                    f"(tile {m.key}{list(idx)} written by {nodes[writer[k]]} and {n})")
                writer[k] = n.nid
            for (m, idx) in n.reads:
                readers.setdefault(_tile_key(m, idx), []).append(n.nid)
        for n in nodes:
            seen = set()
            for (m, idx) in n.writes:
                for c in readers.get(_tile_key(m, idx), ()):
                    if c not in seen:
                        seen.add(c)
                        n.children.append(c)
            seen = set()
            for (m, idx) in n.reads:
                p = writer.get(_tile_key(m, idx))
                if p is not None and p not in seen:
                    seen.add(p)
                    n.parents.append(p)
        self._nodes = nodes
        self._by_key = {n.key: n.nid for n in nodes}
        self._writer = writer
        self._readers = readers
        self.expand_time = time.time() - t0

    @property
    def nodes(self) -> List[ExpandedNode]:
        if self._nodes is None:
            self._expand()
        return self._nodes

    def node(self, expr_idx, var_values) -> ExpandedNode:
        nodes = self.nodes
        nid = self._by_key.get((int(expr_idx), _freeze(var_values)))
        if nid is None:
            raise KeyError(f"({expr_idx}, {var_values}) is not a node of {self.fdef.name}")
        return nodes[nid]

    def num_readers(self, matrix, idx) -> int:
        self.nodes
        return len(self._readers.get(_tile_key(matrix, idx), ()))

    def writer_of(self, matrix, idx) -> Optional[ExpandedNode]:
        nodes = self.nodes
        w = self._writer.get(_tile_key(matrix, idx))
        return None if w is None else nodes[w]

    # ------------------------------------------------------------------ reference API (compiler.py:49-69)
    def _reads_only(self, call: RemoteCallAbstract, names) -> bool:
        return all(a.matrix_name in names for a in call.args if isinstance(a, IndexExpr))

    def _writes_to(self, call: RemoteCallAbstract, names) -> bool:
        return any(o.matrix_name in names for o in call.output)

    @property
    def starters(self) -> List[Node]:
        """All instances of calls that read only input matrices (compiler.py:709-719)."""
        if self._starters is None:
            ok = {i for i, c in self.remote_calls.items() if self._reads_only(c, set(self.inputs))}
            self._starters = self.enumerate_instances(ok)
        return self._starters

    @property
    def num_terminators(self) -> int:
        """Number of instances of calls that write an output matrix (compiler.py:721-732)."""
        if self._num_terminators is None:
            ok = {i for i, c in self.remote_calls.items() if self._writes_to(c, set(self.outputs))}
            self._num_terminators = len(self.enumerate_instances(ok))
        return self._num_terminators

    def find_children(self, i, value_map) -> List[Node]:
        n = self.node(i, value_map)
        return [self.nodes[c].ref for c in n.children]

    def find_parents(self, i, value_map) -> List[Node]:
        n = self.node(i, value_map)
        return [self.nodes[p].ref for p in n.parents]

    def is_terminator(self, i) -> bool:
        return self._writes_to(self.remote_calls[int(i)], set(self.outputs))

    def eval_expr(self, i, value_map):
        """Node → InstructionBlock [RemoteRead..., RemoteCall, RemoteWrite...] (compiler.py:146-180)."""
        from . import lambdapack as lp
        n = self.node(i, value_map)
        reads = [lp.RemoteRead(0, m, *idx) for (m, idx) in n.reads]
        argv = []
        for kind, j in n.arg_layout:
            argv.append(reads[j] if kind == "read" else n.scalars[j])
        symbols = [str(k) for k in range(len(argv))]
        call = lp.RemoteCall(0, n.call.compute, argv, len(n.writes), symbols, **(n.call.kwargs or {}))
        writes = [lp.RemoteWrite(k + len(argv), m, call.results, k, *idx) for k, (m, idx) in enumerate(n.writes)]
        return lp.InstructionBlock(reads + [call] + writes)

    # dict-like access the reference's module-level helpers expect (program[p_idx], program.keys())
    def keys(self):
        return self.remote_calls.keys()

    def __getitem__(self, i):
        return self.remote_calls[i]

    def __len__(self):
        return len(self.remote_calls)


# --------------------------------------------------------------------------- module-level reference API
def lpcompile(function, namespace=None):
    """Parse once, bind later: ``lpcompile(CHOLESKY)(O, I, S, nb, 0)`` (reference compiler.py:25-38)."""
    fdef = frontend.parse(function, namespace)

    def f(*args, **kwargs):
        if kwargs:
            raise exceptions.LambdaPackBackendGenerationException("keyword arguments are not supported")
        return CompiledLambdaPackProgram(fdef, args, inputs=[], outputs=[])
    f.fdef = fdef
    return f


def lpcompile_for_execution(function, inputs, outputs, namespace=None):
    """Reference compiler.py:40-47."""
    fdef = frontend.parse(function, namespace)

    def f(*args, **kwargs):
        if kwargs:
            raise exceptions.LambdaPackBackendGenerationException("keyword arguments are not supported")
        return CompiledLambdaPackProgram(fdef, args, inputs=inputs, outputs=outputs)
    f.fdef = fdef
    return f


def find_starters(program: CompiledLambdaPackProgram, input_matrices) -> List[Node]:
    ok = {i for i, c in program.remote_calls.items() if program._reads_only(c, set(input_matrices))}
    return program.enumerate_instances(ok)


def find_terminators(program: CompiledLambdaPackProgram, output_matrices) -> List[Node]:
    ok = {i for i, c in program.remote_calls.items() if program._writes_to(c, set(output_matrices))}
    return program.enumerate_instances(ok)


def find_children(program: CompiledLambdaPackProgram, idx, value_map) -> List[Node]:
    return program.find_children(idx, value_map)


def find_parents(program: CompiledLambdaPackProgram, idx, value_map) -> List[Node]:
    return program.find_parents(idx, value_map)


def walk_program(program: CompiledLambdaPackProgram) -> List[Node]:
    """All nodes of the program (reference compiler.py:780-791)."""
    return [n.ref for n in program.nodes]


def eval_remote_call(program: CompiledLambdaPackProgram, idx, value_map):
    return program.eval_expr(idx, value_map)
