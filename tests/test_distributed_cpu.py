"""Host-side logic of the multi-GPU path on CPU: world_size 2 and 4 over gloo (tile ownership, SPMD puts, collective
gather, and the NVLink transfer plan derived from the DAG)."""
import os
import socket
import subprocess
import sys

import pytest

from numpywren_b200 import parallel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4])
def test_gloo_world(world):
    port = free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), NPW_B200_DEVICE="cpu", OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_dist_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{out[-3000:]}"
    assert "DIST_OK %d" % world in outs[0]


@pytest.mark.parametrize("world", [2, 4])
def test_engine_over_gloo(world):
    """The multi-rank engine itself (owner computes, transfer plan, tile exchange, failure agreement) on the host: CUDA
    objects replaced by inert stand-ins, kernels by the NumPy C-ABI double, tiles shipped by gloo isend/irecv."""
    port = free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), NPW_B200_DEVICE="cpu", OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_dist_engine_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=400)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{out[-3000:]}"
    assert "DIST_ENGINE_OK %d" % world in outs[0]


def test_grid_shapes_and_ownership():
    assert parallel.factor_grid(1) == (1, 1)
    assert parallel.factor_grid(2) == (1, 2)
    assert parallel.factor_grid(4) == (2, 2)
    assert parallel.factor_grid(8) == (2, 4)
    assert parallel.factor_grid(6) == (2, 3)

    class M:
        def true_block_idx(self, *idx):
            return idx
    g = parallel.ProcessGrid(8, 3)
    assert g.owner(M(), (0, 0)) == 0 and g.owner(M(), (1, 0)) == 4 and g.owner(M(), (1, 3)) == 7
    assert g.owner(M(), (0, 4)) == 4 and g.owner(M(), (1, 5)) == 1   # process row rotates every Q block columns
    assert g.owner(M(), (5, 2, 6)) == g.owner(M(), (2, 6))          # SSA version axis does not move a tile
    assert g.owner(M(), (3,)) == (3 % 2) * 4
    counts = [0] * 8
    for j in range(32):
        for k in range(j + 1):
            counts[g.owner(M(), (j, k))] += 1
    assert max(counts) - min(counts) <= 16 and min(counts) >= 56      # block-cyclic balances the lower triangle
    # ... and the rotation balances the WORK (tile (j,k) receives k updates) to ~1 %
    work = [0] * 8
    for j in range(32):
        for k in range(j + 1):
            work[g.owner(M(), (j, k))] += k + 1
    assert max(work) / (sum(work) / 8) < 1.03
    with pytest.raises(ValueError):
        parallel.ProcessGrid(8, 0, shape=(3, 3))


def test_tsqr_merges_are_dealt_over_all_ranks():
    """alg_wrappers._place_tsqr_tree: leaf j on rank j mod world, merge k of every level on rank k mod world (with the plain
    row-block map every merge of a 2-rank run landed on rank 0: profiles/r02h timeline)."""
    from numpywren_b200 import alg_wrappers
    from numpywren_b200.matrix import BigMatrix
    for world in (2, 4, 8):
        grid = parallel.ProcessGrid(world, 0)
        parallel.set_grid(grid)
        try:
            X = BigMatrix(f"place_tsqr_{world}", shape=(64 * 8, 8), shard_sizes=(8, 8), device="cpu")
            program, meta = alg_wrappers.tsqr(X)
            plan = parallel.TransferPlan(program.program, grid)
            per_level = {}
            for n in program.program.nodes:
                lvl = int(n.var_values.get("level", -1)) + 1
                per_level.setdefault(lvl, [0] * world)[plan.exec_rank[n.nid]] += 1
            assert per_level[0] == [64 // world] * world                      # leaves
            for lvl, counts in per_level.items():
                total = sum(counts)
                assert max(counts) <= -(-total // world), (world, lvl, counts)   # ceil(total / world): round-robin
            # all three outputs of a node share an owner (the engine runs a node where its first output lives)
            for n in program.program.nodes:
                assert len({grid.owner(m, idx) for (m, idx) in n.writes}) == 1
        finally:
            parallel.set_grid(None)


def test_gemm_plain_block_cyclic_halves_the_inbox():
    """alg_wrappers.place_plain_block_cyclic: GEMM_ACC on a 2x4 grid needs 128 remote tiles per rank with the plain map,
    256 with the rotated (Cholesky) map — 64 GB instead of 131 GB of inbox at N=131072 / tile 8192."""
    from numpywren_b200 import alg_wrappers
    from numpywren_b200.matrix import BigMatrix
    grid = parallel.ProcessGrid(8, 0)
    parallel.set_grid(grid)
    try:
        nb = 16
        A = BigMatrix("plain_gemm_A", shape=(nb * 4, nb * 4), shard_sizes=(4, 4), device="cpu")
        B = BigMatrix("plain_gemm_B", shape=(nb * 4, nb * 4), shard_sizes=(4, 4), device="cpu")
        alg_wrappers.place_plain_block_cyclic(A)
        alg_wrappers.place_plain_block_cyclic(B)
        program, meta = alg_wrappers.gemm_kloop(A, B, out_key="plain_gemm_C")
        plan = parallel.TransferPlan(program.program, grid)
        _, _, max_slots = plan.assign_inbox_slots()
        assert max_slots == 128
        C = meta["outputs"][0]
        owners = [[grid.owner(C, (i, j)) for j in range(nb)] for i in range(nb)]
        assert all(owners[i][j] == (i % 2) * 4 + (j % 4) for i in range(nb) for j in range(nb))
        counts = [sum(row.count(r) for row in owners) for r in range(8)]
        assert counts == [nb * nb // 8] * 8
    finally:
        parallel.set_grid(None)
