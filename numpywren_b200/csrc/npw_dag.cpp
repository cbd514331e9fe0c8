// npw_dag.cpp — native LambdaPACK DAG expander (host only, no CUDA): SURVEY §8 f#4.
//
// The reference answers "who are the children / parents of node (expr, {loop vars})" with sympy at run time
// (compiler.py:269-650: template matching + linsolve + brute-force enumeration, 6-11 ms per node per query).  Here the
// whole program is expanded once: loop nests are enumerated on concrete values, every node records the tiles it reads
// and writes, and the edges come from a hash join  written tile -> readers  (exact for SSA programs).  This file is the
// C++ twin of CompiledLambdaPackProgram._expand (numpywren_b200/compiler.py), which remains the fallback and the
// specification: tests/test_native_dag.py requires both to produce identical nodes, tiles and edges.
//
// Input: the program's loop-nest IR serialised by compiler._serialize_ir as a flat int64 array.
//   statement := FOR    [1, slot, E(lo), E(hi), E(step), B(body)]
//              | ASSIGN [2, slot, E(rhs)]
//              | IF     [3, E(cond), B(body), B(else)]
//              | CALL   [4, expr_idx, n_reads, T*, n_writes, T*]        T := [matrix_id, n_idx, E*]
//   B(block)  := [n_words, statement*]         E(expr) := [n_words, postfix ops]
// Expression values follow Python's semantics for the DSL's subset (frontend.py): int / float, true division,
// floor division and modulo with Python's sign rules, ** with negative exponents, and the exact
// ceiling(log(a)/log(b)) for integer powers (frontend._Log).  Anything outside the subset (overflow, type errors)
// makes the call fail with a message; the caller then falls back to the Python expander, which raises the
// reference-compatible exception.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/npw_dag.h"

namespace {

enum Op : int64_t {
  PUSH_I = 0, PUSH_F = 1, LOAD = 2, ADD = 3, SUB = 4, MUL = 5, DIV = 6, FLOORDIV = 7, MOD = 8, POW = 9, NEG = 10, NOT = 11,
  LT = 12, LE = 13, GT = 14, GE = 15, EQ = 16, NE = 17, AND = 18, OR = 19, CEIL = 20, FLOOR = 21, LOG = 22
};

enum Kind : int { UNSET = 0, INT = 1, FLT = 2, LOGV = 3 };

struct Val {
  int kind = UNSET;
  int64_t i = 0;     // INT value; LOGV: argument if log_int
  double f = 0.0;    // FLT value; LOGV: argument if !log_int
  bool log_int = false;
};

struct Fail {
  std::string msg;
};

inline Val mk_i(int64_t v) { Val x; x.kind = INT; x.i = v; return x; }
inline Val mk_f(double v) { Val x; x.kind = FLT; x.f = v; return x; }

double to_double(const Val& v) {
  switch (v.kind) {
    case INT: return static_cast<double>(v.i);
    case FLT: return v.f;
    case LOGV: return std::log(v.log_int ? static_cast<double>(v.i) : v.f);
    default: throw Fail{"use of an unbound variable"};
  }
}

bool truthy(const Val& v) {
  if (v.kind == INT) return v.i != 0;
  if (v.kind == FLT) return v.f != 0.0;
  if (v.kind == LOGV) return true;
  throw Fail{"use of an unbound variable"};
}

int64_t floordiv_i(int64_t a, int64_t b) {
  if (b == 0) throw Fail{"integer division or modulo by zero"};
  int64_t q = a / b, r = a % b;
  if (r != 0 && ((r < 0) != (b < 0))) --q;
  return q;
}

int64_t mod_i(int64_t a, int64_t b) {
  if (b == 0) throw Fail{"integer division or modulo by zero"};
  int64_t r = a % b;
  if (r != 0 && ((r < 0) != (b < 0))) r += b;
  return r;
}

// frontend._Log._ratio: log(a)/log(b), exact when a is an integer power of b
Val log_ratio(const Val& a, const Val& b) {
  const double approx = to_double(a) / to_double(b);
  if (a.log_int && b.log_int && b.i > 1 && a.i >= 1) {
    int64_t r = 0, p = 1;
    while (p <= a.i / b.i && p * b.i <= a.i) {   // p * b <= a without overflow
      p *= b.i;
      ++r;
    }
    if (p == a.i) return mk_i(r);
    double frac = approx - static_cast<double>(r);
    if (frac < 1e-9) frac = 1e-9;
    if (frac > 1.0 - 1e-9) frac = 1.0 - 1e-9;
    return mk_f(static_cast<double>(r) + frac);
  }
  return mk_f(approx);
}

struct Machine {
  std::vector<Val> slots;
  std::vector<Val> stack;

  Val eval(const int64_t* code, int64_t n) {
    stack.clear();
    for (int64_t pc = 0; pc < n; ++pc) {
      const int64_t op = code[pc];
      switch (op) {
        case PUSH_I: stack.push_back(mk_i(code[++pc])); break;
        case PUSH_F: {
          double d;
          const int64_t bits = code[++pc];
          std::memcpy(&d, &bits, sizeof(d));
          stack.push_back(mk_f(d));
          break;
        }
        case LOAD: {
          const Val& v = slots.at(static_cast<size_t>(code[++pc]));
          if (v.kind == UNSET) throw Fail{"use of an unbound variable"};
          stack.push_back(v);
          break;
        }
        case NEG: case NOT: case CEIL: case FLOOR: case LOG: {
          if (stack.empty()) throw Fail{"malformed expression"};
          Val a = stack.back();
          stack.pop_back();
          if (op == NEG) {
            if (a.kind == INT) stack.push_back(mk_i(-a.i));
            else if (a.kind == FLT) stack.push_back(mk_f(-a.f));
            else throw Fail{"bad operand type for unary -"};
          } else if (op == NOT) {
            stack.push_back(mk_i(truthy(a) ? 0 : 1));
          } else if (op == CEIL) {
            stack.push_back(mk_i(static_cast<int64_t>(std::ceil(to_double(a)))));
          } else if (op == FLOOR) {
            stack.push_back(mk_i(static_cast<int64_t>(std::floor(to_double(a)))));
          } else {  // LOG
            Val x;
            x.kind = LOGV;
            if (a.kind == LOGV) a = mk_f(to_double(a));
            if (a.kind == FLT && std::floor(a.f) == a.f && std::fabs(a.f) < 9.0e18) a = mk_i(static_cast<int64_t>(a.f));
            if (a.kind == INT) {
              if (a.i <= 0) throw Fail{"log of non-positive value"};
              x.log_int = true;
              x.i = a.i;
            } else {
              if (!(a.f > 0.0)) throw Fail{"log of non-positive value"};
              x.log_int = false;
              x.f = a.f;
            }
            stack.push_back(x);
          }
          break;
        }
        default: {
          if (stack.size() < 2) throw Fail{"malformed expression"};
          Val b = stack.back();
          stack.pop_back();
          Val a = stack.back();
          stack.pop_back();
          stack.push_back(binary(op, a, b));
        }
      }
    }
    if (stack.size() != 1) throw Fail{"malformed expression"};
    Val r = stack.back();
    if (r.kind == LOGV) r = mk_f(to_double(r));   // Expr.eval: a bare log decays to a float
    return r;
  }

  static Val binary(int64_t op, const Val& a, const Val& b) {
    const bool ints = a.kind == INT && b.kind == INT;
    switch (op) {
      case ADD: case SUB: case MUL: {
        if (ints) {
          int64_t r;
          bool ovf = op == ADD ? __builtin_add_overflow(a.i, b.i, &r)
                   : op == SUB ? __builtin_sub_overflow(a.i, b.i, &r) : __builtin_mul_overflow(a.i, b.i, &r);
          if (ovf) throw Fail{"integer overflow (beyond int64)"};
          return mk_i(r);
        }
        const double x = to_double(a), y = to_double(b);
        return mk_f(op == ADD ? x + y : op == SUB ? x - y : x * y);
      }
      case DIV: {
        if (a.kind == LOGV && b.kind == LOGV) return log_ratio(a, b);
        const double y = to_double(b);
        if (y == 0.0) throw Fail{"division by zero"};
        return mk_f(to_double(a) / y);
      }
      case FLOORDIV: {
        if (ints) return mk_i(floordiv_i(a.i, b.i));
        if (a.kind == LOGV || b.kind == LOGV) throw Fail{"unsupported operand type for //"};
        if (b.f == 0.0 && b.kind == FLT) throw Fail{"float floor division by zero"};
        const double y = to_double(b);
        if (y == 0.0) throw Fail{"float floor division by zero"};
        return mk_f(std::floor(to_double(a) / y));
      }
      case MOD: {
        if (ints) return mk_i(mod_i(a.i, b.i));
        if (a.kind == LOGV || b.kind == LOGV) throw Fail{"unsupported operand type for %"};
        const double x = to_double(a), y = to_double(b);
        if (y == 0.0) throw Fail{"float modulo"};
        double r = std::fmod(x, y);
        if (r != 0.0 && ((r < 0.0) != (y < 0.0))) r += y;
        return mk_f(r);
      }
      case POW: {
        if (a.kind == LOGV || b.kind == LOGV) throw Fail{"unsupported operand type for **"};
        if (ints && b.i >= 0) {
          int64_t r = 1, base = a.i, e = b.i;
          while (e > 0) {
            if (e & 1) {
              if (__builtin_mul_overflow(r, base, &r)) throw Fail{"integer overflow (beyond int64)"};
            }
            e >>= 1;
            if (e > 0 && __builtin_mul_overflow(base, base, &base)) throw Fail{"integer overflow (beyond int64)"};
          }
          return mk_i(r);
        }
        const double x = to_double(a), y = to_double(b);
        if (x == 0.0 && y < 0.0) throw Fail{"0.0 cannot be raised to a negative power"};
        if (x < 0.0 && std::floor(y) != y) throw Fail{"complex result of **"};
        return mk_f(std::pow(x, y));
      }
      case LT: case LE: case GT: case GE: case EQ: case NE: {
        if (a.kind == LOGV || b.kind == LOGV) throw Fail{"comparison of a symbolic log"};
        bool r;
        if (ints) {
          r = op == LT ? a.i < b.i : op == LE ? a.i <= b.i : op == GT ? a.i > b.i : op == GE ? a.i >= b.i
            : op == EQ ? a.i == b.i : a.i != b.i;
        } else {
          const double x = to_double(a), y = to_double(b);
          r = op == LT ? x < y : op == LE ? x <= y : op == GT ? x > y : op == GE ? x >= y : op == EQ ? x == y : x != y;
        }
        return mk_i(r ? 1 : 0);
      }
      case AND: return truthy(a) ? b : a;
      case OR: return truthy(a) ? a : b;
      default: throw Fail{"unknown opcode"};
    }
  }
};

int64_t as_int(const Val& v, const char* what) {   // Python int(x): truncation towards zero
  if (v.kind == INT) return v.i;
  if (v.kind == FLT) {
    if (!std::isfinite(v.f) || std::fabs(v.f) > 9.0e18) throw Fail{std::string("cannot convert ") + what + " to an integer"};
    return static_cast<int64_t>(v.f);
  }
  throw Fail{std::string("cannot convert ") + what + " to an integer"};
}

int64_t as_index(const Val& v) {   // a block index must be integer-valued
  if (v.kind == INT) return v.i;
  if (v.kind == FLT) {
    if (std::floor(v.f) != v.f || std::fabs(v.f) > 9.0e18) throw Fail{"non-integer block index"};
    return static_cast<int64_t>(v.f);
  }
  throw Fail{"non-integer block index"};
}

struct VecHash {
  size_t operator()(const std::vector<int64_t>& v) const {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (int64_t x : v) {
      h ^= static_cast<uint64_t>(x) + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
      h *= 0xBF58476D1CE4E5B9ull;
    }
    return static_cast<size_t>(h ^ (h >> 31));
  }
};

}  // namespace

struct npw_dag {
  Machine vm;
  int64_t max_nodes = 0;
  // nodes
  std::vector<int32_t> node_expr;
  std::vector<int64_t> var_off{0}, var_slot, var_val;
  std::vector<int64_t> r_off{0}, r_tile, w_off{0}, w_tile;
  // tiles
  std::unordered_map<std::vector<int64_t>, int64_t, VecHash> tile_id;   // key = [matrix_id, idx...]
  std::vector<int64_t> t_matrix, t_idx_off{0}, t_idx, t_writer;
  // edges
  std::vector<int64_t> c_off, c_node, p_off, p_node;
  // loop variables currently in scope (nesting order)
  std::vector<std::pair<int64_t, int64_t>> loop_vars;
  std::vector<int64_t> key;

  int64_t intern(int64_t matrix, const int64_t*& pc) {
    const int64_t n_idx = *pc++;
    key.clear();
    key.push_back(matrix);
    for (int64_t a = 0; a < n_idx; ++a) {
      const int64_t len = *pc++;
      key.push_back(as_index(vm.eval(pc, len)));
      pc += len;
    }
    auto it = tile_id.find(key);
    if (it != tile_id.end()) return it->second;
    const int64_t id = static_cast<int64_t>(t_matrix.size());
    tile_id.emplace(key, id);
    t_matrix.push_back(matrix);
    for (size_t a = 1; a < key.size(); ++a) t_idx.push_back(key[a]);
    t_idx_off.push_back(static_cast<int64_t>(t_idx.size()));
    t_writer.push_back(-1);
    return id;
  }

  void block(const int64_t* pc, const int64_t* end) {
    while (pc < end) {
      const int64_t kind = *pc++;
      if (kind == 1) {            // FOR
        const int64_t slot = *pc++;
        const int64_t l0 = *pc++; const int64_t* e0 = pc; pc += l0;
        const int64_t l1 = *pc++; const int64_t* e1 = pc; pc += l1;
        const int64_t l2 = *pc++; const int64_t* e2 = pc; pc += l2;
        const int64_t lb = *pc++; const int64_t* body = pc; pc += lb;
        const int64_t lo = as_int(vm.eval(e0, l0), "a range bound");
        const int64_t hi = as_int(vm.eval(e1, l1), "a range bound");
        const int64_t st = as_int(vm.eval(e2, l2), "a range step");
        if (st == 0) throw Fail{"range() step must not be zero"};
        // assignments inside the body are scoped to one iteration (the Python expander copies its environment)
        const std::vector<Val> saved = vm.slots;
        loop_vars.emplace_back(slot, 0);
        for (int64_t v = lo; st > 0 ? v < hi : v > hi; v += st) {
          vm.slots = saved;
          vm.slots[static_cast<size_t>(slot)] = mk_i(v);
          loop_vars.back().second = v;
          block(body, body + lb);
        }
        loop_vars.pop_back();
        vm.slots = saved;
      } else if (kind == 2) {     // ASSIGN
        const int64_t slot = *pc++;
        const int64_t len = *pc++;
        vm.slots[static_cast<size_t>(slot)] = vm.eval(pc, len);
        pc += len;
      } else if (kind == 3) {     // IF
        const int64_t lc = *pc++; const int64_t* cond = pc; pc += lc;
        const int64_t lb = *pc++; const int64_t* body = pc; pc += lb;
        const int64_t le = *pc++; const int64_t* els = pc; pc += le;
        const std::vector<Val> saved = vm.slots;
        if (truthy(vm.eval(cond, lc))) block(body, body + lb); else block(els, els + le);
        vm.slots = saved;
      } else if (kind == 4) {     // CALL
        const int64_t expr_idx = *pc++;
        const int64_t nid = static_cast<int64_t>(node_expr.size());
        if (max_nodes > 0 && nid >= max_nodes) throw Fail{"program exceeds the node limit"};
        node_expr.push_back(static_cast<int32_t>(expr_idx));
        for (auto& lv : loop_vars) {
          var_slot.push_back(lv.first);
          var_val.push_back(lv.second);
        }
        var_off.push_back(static_cast<int64_t>(var_slot.size()));
        const int64_t n_reads = *pc++;
        for (int64_t a = 0; a < n_reads; ++a) {
          const int64_t matrix = *pc++;
          r_tile.push_back(intern(matrix, pc));
        }
        r_off.push_back(static_cast<int64_t>(r_tile.size()));
        const int64_t n_writes = *pc++;
        for (int64_t a = 0; a < n_writes; ++a) {
          const int64_t matrix = *pc++;
          const int64_t t = intern(matrix, pc);
          if (t_writer[static_cast<size_t>(t)] >= 0) {
            char buf[160];
            std::snprintf(buf, sizeof(buf), "not SSA: tile %lld written by nodes %lld and %lld", static_cast<long long>(t),
                          static_cast<long long>(t_writer[static_cast<size_t>(t)]), static_cast<long long>(nid));
            throw Fail{buf};
          }
          t_writer[static_cast<size_t>(t)] = nid;
          w_tile.push_back(t);
        }
        w_off.push_back(static_cast<int64_t>(w_tile.size()));
      } else {
        throw Fail{"malformed program"};
      }
    }
  }

  void edges() {
    const int64_t n = static_cast<int64_t>(node_expr.size());
    const int64_t nt = static_cast<int64_t>(t_matrix.size());
    // readers of every tile, in node order (a node that reads a tile twice appears twice, like the Python expander)
    std::vector<int64_t> rd_off(static_cast<size_t>(nt) + 1, 0), rd;
    for (int64_t t : r_tile) ++rd_off[static_cast<size_t>(t) + 1];
    for (int64_t t = 0; t < nt; ++t) rd_off[static_cast<size_t>(t) + 1] += rd_off[static_cast<size_t>(t)];
    rd.resize(r_tile.size());
    std::vector<int64_t> fill(rd_off.begin(), rd_off.end() - 1);
    for (int64_t v = 0; v < n; ++v)
      for (int64_t a = r_off[static_cast<size_t>(v)]; a < r_off[static_cast<size_t>(v) + 1]; ++a)
        rd[static_cast<size_t>(fill[static_cast<size_t>(r_tile[static_cast<size_t>(a)])]++)] = v;
    std::vector<int64_t> stamp(static_cast<size_t>(n), -1);
    c_off.assign(1, 0);
    for (int64_t v = 0; v < n; ++v) {
      for (int64_t a = w_off[static_cast<size_t>(v)]; a < w_off[static_cast<size_t>(v) + 1]; ++a) {
        const int64_t t = w_tile[static_cast<size_t>(a)];
        for (int64_t q = rd_off[static_cast<size_t>(t)]; q < rd_off[static_cast<size_t>(t) + 1]; ++q) {
          const int64_t c = rd[static_cast<size_t>(q)];
          if (stamp[static_cast<size_t>(c)] != v) {
            stamp[static_cast<size_t>(c)] = v;
            c_node.push_back(c);
          }
        }
      }
      c_off.push_back(static_cast<int64_t>(c_node.size()));
    }
    std::fill(stamp.begin(), stamp.end(), -1);
    p_off.assign(1, 0);
    for (int64_t v = 0; v < n; ++v) {
      for (int64_t a = r_off[static_cast<size_t>(v)]; a < r_off[static_cast<size_t>(v) + 1]; ++a) {
        const int64_t p = t_writer[static_cast<size_t>(r_tile[static_cast<size_t>(a)])];
        if (p >= 0 && stamp[static_cast<size_t>(p)] != v) {
          stamp[static_cast<size_t>(p)] = v;
          p_node.push_back(p);
        }
      }
      p_off.push_back(static_cast<int64_t>(p_node.size()));
    }
  }
};

extern "C" {

npw_dag* npw_dag_expand(const int64_t* code, int64_t code_len, int32_t n_slots, const int8_t* slot_kind,
                        const int64_t* slot_int, const double* slot_float, int64_t max_nodes, char* err, int32_t err_len) {
  if (err && err_len > 0) err[0] = 0;
  if (!code || code_len < 0 || n_slots < 0) {
    if (err && err_len > 0) std::snprintf(err, static_cast<size_t>(err_len), "bad argument");
    return nullptr;
  }
  npw_dag* d = new npw_dag();
  try {
    d->max_nodes = max_nodes;
    d->vm.slots.resize(static_cast<size_t>(n_slots));
    for (int32_t s = 0; s < n_slots; ++s) {
      if (slot_kind[s] == INT) d->vm.slots[static_cast<size_t>(s)] = mk_i(slot_int[s]);
      else if (slot_kind[s] == FLT) d->vm.slots[static_cast<size_t>(s)] = mk_f(slot_float[s]);
    }
    d->block(code, code + code_len);
    d->edges();
  } catch (const Fail& f) {
    if (err && err_len > 0) std::snprintf(err, static_cast<size_t>(err_len), "%s", f.msg.c_str());
    delete d;
    return nullptr;
  } catch (const std::exception& e) {
    if (err && err_len > 0) std::snprintf(err, static_cast<size_t>(err_len), "%s", e.what());
    delete d;
    return nullptr;
  }
  return d;
}

void npw_dag_arrays(const npw_dag* d, npw_dag_view* v) {
  v->n_nodes = static_cast<int64_t>(d->node_expr.size());
  v->n_tiles = static_cast<int64_t>(d->t_matrix.size());
  v->node_expr = d->node_expr.data();
  v->var_off = d->var_off.data(); v->var_slot = d->var_slot.data(); v->var_val = d->var_val.data();
  v->read_off = d->r_off.data(); v->read_tile = d->r_tile.data();
  v->write_off = d->w_off.data(); v->write_tile = d->w_tile.data();
  v->tile_matrix = d->t_matrix.data(); v->tile_idx_off = d->t_idx_off.data(); v->tile_idx = d->t_idx.data();
  v->tile_writer = d->t_writer.data();
  v->child_off = d->c_off.data(); v->child = d->c_node.data();
  v->parent_off = d->p_off.data(); v->parent = d->p_node.data();
}

void npw_dag_free(npw_dag* d) { delete d; }

int npw_dag_abi_version(void) { return 1; }

}  // extern "C"
