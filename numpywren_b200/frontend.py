"""LambdaPACK front end: Python function source → a small loop-nest IR.

Accepts the reference DSL (grammar in reference frontend.py:47-72): ``for v in range(...)``
nests, static ``if/else``, scalar assignments, and remote calls
``M[i, j], ... = kernel(A[i, k], B[k, j], 2.0)`` whose kernel is looked up *by name*
(reference frontend.py:343 does ``eval(name)`` inside a namespace that star-imported
kernels.py; here the namespace is ``numpywren_b200.kernels`` plus anything the caller adds).

Unlike the reference there is no sympy: index expressions are compiled to Python code
objects and evaluated on concrete loop values when the program is expanded (compiler.py).
``ceiling(log(a)/log(b))`` — the only non-affine form the reference programs use
(algs.py:22,34,186,252) — is evaluated exactly for integer powers, like sympy would.
"""
from __future__ import annotations

import ast
import inspect
import math
import textwrap
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Tuple

from . import exceptions

M_FUNCS = ("ceiling", "floor", "log")
_VALID_BINOPS = (ast.Add, ast.Sub, ast.Mult, ast.Div, ast.Mod, ast.Pow, ast.FloorDiv)
_VALID_CMPOPS = (ast.Eq, ast.NotEq, ast.Lt, ast.Gt, ast.LtE, ast.GtE)


# --------------------------------------------------------------------------- exact log/ceiling
class _Log:
    """ln(x) kept symbolic so that log(a)/log(b) is exact when a is an integer power of b."""
    __slots__ = ("x",)

    def __init__(self, x):
        self.x = x

    def __float__(self):
        return math.log(self.x)

    def _ratio(self, other):
        a, b = self.x, other.x
        approx = math.log(a) / math.log(b)
        if isinstance(a, int) and isinstance(b, int) and b > 1 and a >= 1:
            # exact integer part: largest r with b**r <= a; the float quotient only supplies the fraction
            r, p = 0, 1
            while p * b <= a:
                p *= b
                r += 1
            if p == a:
                return r
            return r + min(max(approx - r, 1e-9), 1.0 - 1e-9)
        return approx

    def __truediv__(self, other):
        if isinstance(other, _Log):
            return self._ratio(other)
        return float(self) / other

    def __rtruediv__(self, other):
        return other / float(self)

    def __mul__(self, other):
        return float(self) * float(other)

    __rmul__ = __mul__

    def __add__(self, other):
        return float(self) + float(other)

    __radd__ = __add__

    def __sub__(self, other):
        return float(self) - float(other)

    def __rsub__(self, other):
        return float(other) - float(self)


def _log(x):
    if isinstance(x, _Log):
        x = float(x)
    if isinstance(x, float) and x.is_integer():
        x = int(x)
    if x <= 0:
        raise exceptions.LambdaPackParsingException(f"log of non-positive value {x}")
    return _Log(x)


def _ceiling(x):
    return int(math.ceil(float(x)))


def _floor(x):
    return int(math.floor(float(x)))


EXPR_GLOBALS = {"__builtins__": {}, "ceiling": _ceiling, "floor": _floor, "log": _log, "True": True, "False": False}


# --------------------------------------------------------------------------- IR
@dataclass
class Expr:
    """A scalar DSL expression: source text + compiled code object."""
    src: str
    code: Any
    const: Optional[Any] = None  # literal value when the expression is a bare number

    def eval(self, env: Dict[str, Any]):
        v = eval(self.code, EXPR_GLOBALS, env)
        if isinstance(v, _Log):
            v = float(v)
        return v


@dataclass
class IndexExpr:
    matrix_name: str
    indices: List[Expr]


@dataclass
class RemoteCallAbstract:
    compute: Callable
    compute_name: str
    output: List[IndexExpr]
    args: List[Any]            # IndexExpr | Expr
    kwargs: Dict[str, Any]
    lineno: int = 0


@dataclass
class Assign:
    name: str
    rhs: Expr


@dataclass
class For:
    var: str
    min: Expr
    max: Expr
    step: Expr
    body: List[Any]


@dataclass
class If:
    cond: Expr
    body: List[Any]
    elseBody: List[Any] = field(default_factory=list)


@dataclass
class FuncDef:
    name: str
    args: List[str]
    arg_types: List[Any]
    body: List[Any]
    num_calls: int = 0


# --------------------------------------------------------------------------- parser
class LambdaPackParse:
    """ast → IR.  Mirrors the accept/reject behaviour of reference frontend.LambdaPackParse (:223-479)."""

    def __init__(self, namespace: Optional[Dict[str, Any]] = None):
        from . import kernels as _kernels
        self.namespace = {k: getattr(_kernels, k) for k in dir(_kernels) if not k.startswith("_")}
        if namespace:
            self.namespace.update(namespace)
        self.decls: Dict[str, str] = {}
        self.num_calls = 0

    # ---- expressions
    def _check_expr(self, node):
        for sub in ast.walk(node):
            if isinstance(sub, ast.BinOp):
                if not isinstance(sub.op, _VALID_BINOPS):
                    raise NotImplementedError("Unsupported BinOp {0}".format(type(sub.op).__name__))
            elif isinstance(sub, ast.Compare):
                if len(sub.ops) != 1 or len(sub.comparators) != 1:
                    raise NotImplementedError("Only single op compares supported")
                if not isinstance(sub.ops[0], _VALID_CMPOPS):
                    raise NotImplementedError("Unsupported CmpOp {0}".format(type(sub.ops[0]).__name__))
            elif isinstance(sub, ast.UnaryOp):
                if not isinstance(sub.op, (ast.USub, ast.Not)):
                    raise NotImplementedError("Unsupported unary operation {0}".format(type(sub.op).__name__))
            elif isinstance(sub, ast.Call):
                if not (isinstance(sub.func, ast.Name) and sub.func.id in M_FUNCS):
                    raise exceptions.LambdaPackParsingException("unsupported function in expression")
                if len(sub.args) != 1:
                    raise exceptions.LambdaPackParsingException("m_func calls must single argument")
            elif isinstance(sub, ast.Constant):
                if isinstance(sub.value, str):
                    raise NotImplementedError("Stings not supported")
                if not isinstance(sub.value, (int, float, bool)):
                    raise NotImplementedError("Only Integers and Floats supported")
            elif isinstance(sub, (ast.Name, ast.Load, ast.BoolOp, ast.And, ast.Or, ast.operator, ast.cmpop, ast.unaryop,
                                  ast.Expression)):
                pass
            else:
                raise NotImplementedError("Unsupported expression node {0}".format(type(sub).__name__))

    def expr(self, node) -> Expr:
        self._check_expr(node)
        src = ast.unparse(node)
        code = compile(ast.fix_missing_locations(ast.Expression(body=node)), "<lambdapack>", "eval")
        const = node.value if isinstance(node, ast.Constant) else None
        if isinstance(node, ast.UnaryOp) and isinstance(node.op, ast.USub) and isinstance(node.operand, ast.Constant):
            const = -node.operand.value
        return Expr(src, code, const)

    def index_expr(self, node) -> IndexExpr:
        if not isinstance(node, ast.Subscript) or not isinstance(node.value, ast.Name):
            raise exceptions.LambdaPackParsingException("expected an index expression M[i, ...]")
        sl = node.slice
        elts = sl.elts if isinstance(sl, ast.Tuple) else [sl]
        for e in elts:
            if isinstance(e, ast.Slice):
                raise NotImplementedError("slices are not supported in index expressions")
        return IndexExpr(node.value.id, [self.expr(e) for e in elts])

    # ---- statements
    def stmt(self, node, in_branch=None):
        if isinstance(node, ast.For):
            return self.for_(node)
        if isinstance(node, ast.If):
            return self.if_(node)
        if isinstance(node, ast.Assign):
            return self.assign(node, in_branch)
        if isinstance(node, ast.Expr) and isinstance(node.value, ast.Constant):
            return None  # docstring / bare literal
        if isinstance(node, ast.Return):
            raise exceptions.LambdaPackParsingException(
                "returns forbidden in lambdapack, pass in outputs as function arguments")
        if isinstance(node, ast.Pass):
            return None
        raise NotImplementedError("Unsupported statement {0}".format(type(node).__name__))

    def block(self, nodes, in_branch=None):
        out = []
        for n in nodes:
            s = self.stmt(n, in_branch)
            if s is not None:
                out.append(s)
        return out

    def for_(self, node: ast.For) -> For:
        it = node.iter
        if not (isinstance(it, ast.Call) and isinstance(it.func, ast.Name) and it.func.id == "range"):
            raise NotImplementedError("Only for(x in range(...)) loops allowed")
        if not isinstance(node.target, ast.Name):
            raise NotImplementedError("loop target must be a name")
        zero = self.expr(ast.Constant(0))
        one = self.expr(ast.Constant(1))
        if len(it.args) == 1:
            lo, hi, st = zero, self.expr(it.args[0]), one
        elif len(it.args) == 2:
            lo, hi, st = self.expr(it.args[0]), self.expr(it.args[1]), one
        elif len(it.args) == 3:
            lo, hi, st = self.expr(it.args[0]), self.expr(it.args[1]), self.expr(it.args[2])
        else:
            raise NotImplementedError("range() takes 1 to 3 arguments")
        body = self.block(node.body)
        self.decls[node.target.id] = "loop"
        return For(node.target.id, lo, hi, st, body)

    def if_(self, node: ast.If) -> If:
        cond = self.expr(node.test)
        outer = dict(self.decls)
        body = self.block(node.body, in_branch="if")
        declared_if = {s.name for s in body if isinstance(s, Assign)}
        self.decls = dict(outer)
        else_body = self.block(node.orelse, in_branch="else")
        declared_else = {s.name for s in else_body if isinstance(s, Assign)}
        for name in declared_else - declared_if:
            raise exceptions.LambdaPackParsingException("Variable {0} declared in else but not in if".format(name))
        if node.orelse and declared_if != declared_else:
            raise exceptions.LambdaPackParsingException("if/else didn't have symmetric pair of declarations")
        self.decls = dict(outer)
        for name in declared_if:
            self.decls[name] = "var"
        return If(cond, body, else_body)

    def assign(self, node: ast.Assign, in_branch):
        rhs = node.value
        is_remote = isinstance(rhs, ast.Call) and isinstance(rhs.func, ast.Name) and rhs.func.id not in M_FUNCS
        if is_remote:
            name = rhs.func.id
            fn = self.namespace.get(name)
            if fn is None or not callable(fn):
                raise Exception("unsupported function {0}".format(name))
            if len(node.targets) != 1:
                raise NotImplementedError("chained assignment is not supported")
            tgt = node.targets[0]
            outs = tgt.elts if isinstance(tgt, ast.Tuple) else [tgt]
            outputs = [self.index_expr(o) for o in outs]
            args = []
            for a in rhs.args:
                if isinstance(a, ast.Starred):
                    raise NotImplementedError("starred arguments are not supported")
                args.append(self.index_expr(a) if isinstance(a, ast.Subscript) else self.expr(a))
            # the reference evaluates but then drops keyword arguments (frontend.py:331,346): keep that behaviour
            call = RemoteCallAbstract(fn, name, outputs, args, {}, getattr(node, "lineno", 0))
            self.num_calls += 1
            return call
        if len(node.targets) != 1 or not isinstance(node.targets[0], ast.Name):
            raise NotImplementedError("Multiple targets only supported for RemoteOps")
        name = node.targets[0].id
        if name in self.decls:
            raise exceptions.LambdaPackParsingException("multiple variable declarations forbidden")
        self.decls[name] = "var"
        return Assign(name, self.expr(rhs))

    def visit(self, func: ast.FunctionDef) -> FuncDef:
        args = [a.arg for a in func.args.args]
        if len(set(args)) != len(args):
            raise exceptions.LambdaPackParsingException("No repeat arguments allowed")
        types = []
        for a in func.args.args:
            ann = a.annotation
            types.append(ann.id if isinstance(ann, ast.Name) else (ann.attr if isinstance(ann, ast.Attribute) else None))
        for a in args:
            self.decls[a] = "arg"
        body = self.block(func.body)
        return FuncDef(func.name, args, types, body, self.num_calls)


def parse(function, namespace=None) -> FuncDef:
    """Parse a LambdaPACK program given as a Python function (or its source text)."""
    src = function if isinstance(function, str) else inspect.getsource(function)
    tree = ast.parse(textwrap.dedent(src))
    fdef = tree.body[0]
    if not isinstance(fdef, ast.FunctionDef):
        raise exceptions.LambdaPackParsingException("expected a function definition")
    return LambdaPackParse(namespace).visit(fdef)
