/*
 * npw_dag.h — C-ABI of libnpw_dag.so: native expansion of a LambdaPACK program into its tile DAG (host code, no CUDA).
 *
 * Replaces, for programs over plain BigMatrix arguments, what the reference computes symbolically at run time:
 *   compiler.walk_program            (numpywren/compiler.py:780-791)   -> the node list
 *   compiler.find_children / find_parents (compiler.py:595-650)        -> the CSR edge arrays
 *   compiler.eval_remote_call        (compiler.py:146-180)             -> the tiles every node reads / writes
 * Node identity is the reference's: (expr_idx = position of the remote call in source order, {loop variable: value}).
 * The Python host side (numpywren_b200/compiler.py) serialises its loop-nest IR, calls npw_dag_expand once per
 * program and wraps the arrays; without the library it runs its own (identical, slower) expander.
 *
 * All arrays are owned by the handle and stay valid until npw_dag_free.  Offsets are CSR: entries of node v are
 * [off[v], off[v+1]).
 */
#ifndef NPW_DAG_H
#define NPW_DAG_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct npw_dag npw_dag;

typedef struct npw_dag_view {
  int64_t n_nodes, n_tiles;
  const int32_t* node_expr;       /* [n_nodes] remote-call index of the node                                   */
  const int64_t* var_off;         /* [n_nodes+1] loop variables in scope, outermost first:                     */
  const int64_t* var_slot;        /*   variable slot (index into the caller's name table)                      */
  const int64_t* var_val;         /*   and its value                                                           */
  const int64_t* read_off;        /* [n_nodes+1] tiles read, in argument order                                 */
  const int64_t* read_tile;
  const int64_t* write_off;       /* [n_nodes+1] tiles written, in output order                                */
  const int64_t* write_tile;
  const int64_t* tile_matrix;     /* [n_tiles] matrix id of the tile (caller's table)                          */
  const int64_t* tile_idx_off;    /* [n_tiles+1] block index of the tile                                       */
  const int64_t* tile_idx;
  const int64_t* tile_writer;     /* [n_tiles] node that writes the tile, -1 if none (program input / default) */
  const int64_t* child_off;       /* [n_nodes+1] find_children: readers of the node's outputs, deduplicated    */
  const int64_t* child;
  const int64_t* parent_off;      /* [n_nodes+1] find_parents: writers of the node's inputs, deduplicated      */
  const int64_t* parent;
} npw_dag_view;

/* Expand a serialised program (format: csrc/npw_dag.cpp header).  `slot_kind[s]` is 0 (unbound), 1 (int: slot_int[s])
 * or 2 (float: slot_float[s]) for the program's scalar arguments.  max_nodes > 0 bounds the expansion.  Returns NULL on
 * failure with a message in err (expression outside the supported subset, non-SSA program, node limit). */
npw_dag* npw_dag_expand(const int64_t* code, int64_t code_len, int32_t n_slots, const int8_t* slot_kind,
                        const int64_t* slot_int, const double* slot_float, int64_t max_nodes, char* err, int32_t err_len);
void npw_dag_arrays(const npw_dag* dag, npw_dag_view* out);
void npw_dag_free(npw_dag* dag);
int npw_dag_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NPW_DAG_H */
