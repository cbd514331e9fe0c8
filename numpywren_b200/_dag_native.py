"""ctypes binding of libnpw_dag.so (include/npw_dag.h): the native LambdaPACK DAG expander.

``expand(program)`` serialises the loop-nest IR of a ``CompiledLambdaPackProgram`` into the flat int64 format documented
in csrc/npw_dag.cpp, runs the C++ expander and returns the node / tile / edge arrays as NumPy arrays.  It returns
``None`` — and the caller runs the Python expander, the specification of the result — when the library is not built, when
the program uses something the native side does not model (scalar kernel arguments, BigMatrixView arguments), or when
the native evaluator reports a condition for which Python must raise its own exception.  This is host-side scheduling
logic, not the GPU compute path: falling back here changes speed only, never results (tests/test_native_dag.py pins the
two expanders to identical output).
"""
from __future__ import annotations

import ast
import ctypes
import os
import struct
from ctypes import POINTER, c_char_p, c_double, c_int8, c_int32, c_int64, c_void_p
from typing import Any, Dict, List, Optional

import numpy as np

from .frontend import Assign, For, If, IndexExpr, RemoteCallAbstract

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libnpw_dag.so")
_lib = None
_load_failed = False

OPS = {"PUSH_I": 0, "PUSH_F": 1, "LOAD": 2, "ADD": 3, "SUB": 4, "MUL": 5, "DIV": 6, "FLOORDIV": 7, "MOD": 8, "POW": 9,
       "NEG": 10, "NOT": 11, "LT": 12, "LE": 13, "GT": 14, "GE": 15, "EQ": 16, "NE": 17, "AND": 18, "OR": 19, "CEIL": 20,
       "FLOOR": 21, "LOG": 22}
_BIN = {ast.Add: "ADD", ast.Sub: "SUB", ast.Mult: "MUL", ast.Div: "DIV", ast.FloorDiv: "FLOORDIV", ast.Mod: "MOD",
        ast.Pow: "POW"}
_CMP = {ast.Lt: "LT", ast.LtE: "LE", ast.Gt: "GT", ast.GtE: "GE", ast.Eq: "EQ", ast.NotEq: "NE"}
_FUN = {"ceiling": "CEIL", "floor": "FLOOR", "log": "LOG"}


class Unsupported(Exception):
    """The program needs the Python expander."""


class _View(ctypes.Structure):
    _fields_ = [("n_nodes", c_int64), ("n_tiles", c_int64), ("node_expr", POINTER(c_int32)),
                ("var_off", POINTER(c_int64)), ("var_slot", POINTER(c_int64)), ("var_val", POINTER(c_int64)),
                ("read_off", POINTER(c_int64)), ("read_tile", POINTER(c_int64)),
                ("write_off", POINTER(c_int64)), ("write_tile", POINTER(c_int64)),
                ("tile_matrix", POINTER(c_int64)), ("tile_idx_off", POINTER(c_int64)), ("tile_idx", POINTER(c_int64)),
                ("tile_writer", POINTER(c_int64)),
                ("child_off", POINTER(c_int64)), ("child", POINTER(c_int64)),
                ("parent_off", POINTER(c_int64)), ("parent", POINTER(c_int64))]


def load():
    """The library handle, or None when it is not built / cannot be loaded (the Python expander is used then)."""
    global _lib, _load_failed
    if _lib is not None or _load_failed:
        return _lib
    if os.environ.get("NPW_B200_NATIVE_DAG", "1") == "0" or not os.path.exists(_LIB_PATH):
        _load_failed = True
        return None
    try:
        lib = ctypes.CDLL(_LIB_PATH)
        lib.npw_dag_expand.restype = c_void_p
        lib.npw_dag_expand.argtypes = [POINTER(c_int64), c_int64, c_int32, POINTER(c_int8), POINTER(c_int64), POINTER(c_double),
                                       c_int64, c_char_p, c_int32]
        lib.npw_dag_arrays.restype = None
        lib.npw_dag_arrays.argtypes = [c_void_p, POINTER(_View)]
        lib.npw_dag_free.restype = None
        lib.npw_dag_free.argtypes = [c_void_p]
        lib.npw_dag_abi_version.restype = ctypes.c_int
        if lib.npw_dag_abi_version() != 1:
            raise OSError("libnpw_dag.so ABI version mismatch")
        _lib = lib
    except (OSError, AttributeError):
        _load_failed = True
    return _lib


# --------------------------------------------------------------------------------------------- serialisation
class _Serializer:
    def __init__(self, program):
        self.program = program
        self.slots: Dict[str, int] = {}
        self.matrix_ids: Dict[Any, int] = {}
        self.matrices: List[Any] = []
        self.call_idx = {id(c): i for i, c in program.remote_calls.items()}

    def slot(self, name: str) -> int:
        s = self.slots.get(name)
        if s is None:
            s = self.slots[name] = len(self.slots)
        return s

    def matrix(self, name: str) -> int:
        from .matrix import BigMatrixView
        m = self.program._matrix(name)
        if isinstance(m, BigMatrixView) or getattr(m, "transposed", False):
            raise Unsupported("view arguments remap block indices in Python")
        ident = (getattr(m, "bucket", None), getattr(m, "key", id(m)))
        mid = self.matrix_ids.get(ident)
        if mid is None:
            mid = self.matrix_ids[ident] = len(self.matrices)
            self.matrices.append(m)
        return mid

    # ---- expressions → postfix
    def expr(self, e) -> List[int]:
        out: List[int] = []
        self._emit(ast.parse(e.src, mode="eval").body, out)
        return [len(out)] + out

    def _emit(self, node, out):
        if isinstance(node, ast.Constant):
            v = node.value
            if isinstance(v, bool):
                out += [OPS["PUSH_I"], int(v)]
            elif isinstance(v, int):
                if not -2 ** 63 <= v < 2 ** 63:
                    raise Unsupported("integer literal beyond int64")
                out += [OPS["PUSH_I"], v]
            elif isinstance(v, float):
                out += [OPS["PUSH_F"], struct.unpack("<q", struct.pack("<d", v))[0]]
            else:
                raise Unsupported("literal type")
        elif isinstance(node, ast.Name):
            if node.id in ("True", "False"):
                out += [OPS["PUSH_I"], int(node.id == "True")]
            else:
                out += [OPS["LOAD"], self.slot(node.id)]
        elif isinstance(node, ast.BinOp) and type(node.op) in _BIN:
            self._emit(node.left, out)
            self._emit(node.right, out)
            out.append(OPS[_BIN[type(node.op)]])
        elif isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.USub, ast.Not)):
            self._emit(node.operand, out)
            out.append(OPS["NEG" if isinstance(node.op, ast.USub) else "NOT"])
        elif isinstance(node, ast.Compare) and len(node.ops) == 1 and type(node.ops[0]) in _CMP:
            self._emit(node.left, out)
            self._emit(node.comparators[0], out)
            out.append(OPS[_CMP[type(node.ops[0])]])
        elif isinstance(node, ast.BoolOp) and isinstance(node.op, (ast.And, ast.Or)):
            self._emit(node.values[0], out)
            for v in node.values[1:]:
                self._emit(v, out)
                out.append(OPS["AND" if isinstance(node.op, ast.And) else "OR"])
        elif isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id in _FUN and len(node.args) == 1:
            self._emit(node.args[0], out)
            out.append(OPS[_FUN[node.func.id]])
        else:
            raise Unsupported("expression node " + type(node).__name__)

    # ---- statements
    def tile(self, ie: IndexExpr) -> List[int]:
        out = [self.matrix(ie.matrix_name), len(ie.indices)]
        for e in ie.indices:
            out += self.expr(e)
        return out

    def block(self, body) -> List[int]:
        out: List[int] = []
        for s in body:
            if isinstance(s, RemoteCallAbstract):
                reads = [a for a in s.args if isinstance(a, IndexExpr)]
                if len(reads) != len(s.args):
                    raise Unsupported("scalar kernel arguments are evaluated in Python")
                out += [4, self.call_idx[id(s)], len(reads)]
                for a in reads:
                    out += self.tile(a)
                out.append(len(s.output))
                for o in s.output:
                    out += self.tile(o)
            elif isinstance(s, Assign):
                out += [2, self.slot(s.name)] + self.expr(s.rhs)
            elif isinstance(s, For):
                out += [1, self.slot(s.var)] + self.expr(s.min) + self.expr(s.max) + self.expr(s.step) + self.block(s.body)
            elif isinstance(s, If):
                out += [3] + self.expr(s.cond) + self.block(s.body) + self.block(s.elseBody)
            else:
                raise Unsupported("statement " + type(s).__name__)
        return [len(out)] + out


def expand(program, max_nodes: int = 0) -> Optional[Dict[str, Any]]:
    """Native expansion of ``program`` (a CompiledLambdaPackProgram) → dict of NumPy arrays plus the slot-name and
    matrix tables, or None when the Python expander must be used."""
    lib = load()
    if lib is None:
        return None
    ser = _Serializer(program)
    try:
        # scalar arguments first, so that their slots exist before the body refers to them
        scalars = {k: v for k, v in program.scope.items() if not (hasattr(v, "get_block") and hasattr(v, "shard_sizes"))}
        for name in scalars:
            ser.slot(name)
        code = ser.block(program.fdef.body)[1:]
    except Unsupported:
        return None
    n_slots = len(ser.slots)
    kind = np.zeros(n_slots, dtype=np.int8)
    ival = np.zeros(n_slots, dtype=np.int64)
    fval = np.zeros(n_slots, dtype=np.float64)
    for name, v in scalars.items():
        s = ser.slots[name]
        if isinstance(v, bool) or isinstance(v, (int, np.integer)):
            if not -2 ** 63 <= int(v) < 2 ** 63:
                return None
            kind[s], ival[s] = 1, int(v)
        elif isinstance(v, (float, np.floating)):
            kind[s], fval[s] = 2, float(v)
        else:
            return None
    try:
        arr = np.asarray(code, dtype=np.int64)
    except OverflowError:
        return None
    err = ctypes.create_string_buffer(256)
    h = lib.npw_dag_expand(arr.ctypes.data_as(POINTER(c_int64)), arr.size, n_slots, kind.ctypes.data_as(POINTER(c_int8)),
                           ival.ctypes.data_as(POINTER(c_int64)), fval.ctypes.data_as(POINTER(c_double)), int(max_nodes), err,
                           len(err))
    if not h:
        return None        # the Python expander raises the reference-compatible exception for this program
    try:
        v = _View()
        lib.npw_dag_arrays(h, ctypes.byref(v))
        n, nt = int(v.n_nodes), int(v.n_tiles)

        def take(ptr, count, dtype=np.int64):
            if count == 0:
                return np.zeros(0, dtype=dtype)
            return np.ctypeslib.as_array(ptr, shape=(count,)).copy()

        var_off = take(v.var_off, n + 1)
        read_off = take(v.read_off, n + 1)
        write_off = take(v.write_off, n + 1)
        tile_idx_off = take(v.tile_idx_off, nt + 1)
        child_off = take(v.child_off, n + 1)
        parent_off = take(v.parent_off, n + 1)
        out = {"n_nodes": n, "n_tiles": nt, "node_expr": take(v.node_expr, n, np.int32), "var_off": var_off,
               "var_slot": take(v.var_slot, int(var_off[-1])), "var_val": take(v.var_val, int(var_off[-1])),
               "read_off": read_off, "read_tile": take(v.read_tile, int(read_off[-1])),
               "write_off": write_off, "write_tile": take(v.write_tile, int(write_off[-1])),
               "tile_matrix": take(v.tile_matrix, nt), "tile_idx_off": tile_idx_off,
               "tile_idx": take(v.tile_idx, int(tile_idx_off[-1])), "tile_writer": take(v.tile_writer, nt),
               "child_off": child_off, "child": take(v.child, int(child_off[-1])),
               "parent_off": parent_off, "parent": take(v.parent, int(parent_off[-1])),
               "slot_names": [name for name, _ in sorted(ser.slots.items(), key=lambda kv: kv[1])],
               "matrices": ser.matrices}
    finally:
        lib.npw_dag_free(h)
    return out
