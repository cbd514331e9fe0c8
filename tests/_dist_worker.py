"""Worker for tests/test_distributed_cpu.py: run under torchrun-style env (RANK/WORLD_SIZE/MASTER_*), gloo backend."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from numpywren_b200 import algs, compiler, parallel  # noqa: E402
from numpywren_b200.matrix import BigMatrix  # noqa: E402
from numpywren_b200.matrix_init import shard_matrix  # noqa: E402
from numpywren_b200.matrix_utils import constant_zeros  # noqa: E402


def main():
    grid = parallel.init_from_env("gloo")
    world, rank = grid.world, grid.rank
    assert (grid.P, grid.Q) == parallel.factor_grid(world)

    # ---- ownership: SPMD put keeps only owned tiles; numpy() is a collective gather
    X = np.arange(36 * 36, dtype=np.float64).reshape(36, 36)
    A = BigMatrix("dist_A", shape=X.shape, shard_sizes=(8, 8), device="cpu")
    shard_matrix(A, X)
    mine = [b for b in A.block_idxs if grid.is_mine(A, b)]
    assert sorted(A.block_idxs_exist) == sorted(mine)
    counts = [None] * world
    dist.all_gather_object(counts, len(mine))
    assert sum(counts) == len(A.block_idxs) and max(counts) - min(counts) <= 5
    assert np.array_equal(A.numpy(), X)
    other = [b for b in A.block_idxs if not grid.is_mine(A, b)][0]
    try:
        A.get_block(*other)
        raise SystemExit("remote get_block should raise")
    except Exception as e:
        assert "owned by rank" in str(e)
    Z = BigMatrix("dist_Z", shape=(16, 16), shard_sizes=(8, 8), device="cpu", parent_fn=constant_zeros)
    assert not Z.numpy().any()

    # ---- transfer plan of the Cholesky DAG: identical on all ranks, sends and recvs pair up in order
    nb = 6
    O = BigMatrix("dist_O", shape=(nb * 4, nb * 4), shard_sizes=(4, 4), device="cpu")
    I = BigMatrix("dist_I", shape=(nb * 4, nb * 4), shard_sizes=(4, 4), device="cpu")
    S = BigMatrix("dist_S", shape=(nb + 1, nb * 4, nb * 4), shard_sizes=(1, 4, 4), device="cpu")
    prog = compiler.lpcompile_for_execution(algs.CHOLESKY, ["I"], ["O"])(O, I, S, nb, 0)
    plan = parallel.TransferPlan(prog, grid)
    # owner computes: every node runs where its output tile lives, and in-place chains stay on one rank
    for n in prog.nodes:
        assert plan.exec_rank[n.nid] == grid.owner(*n.writes[0])
        if n.call.compute_name == "syrk":
            assert grid.owner(*n.reads[0]) == plan.exec_rank[n.nid]       # S[i,j,k] -> S[i+1,j,k] never moves
        if n.call.compute_name == "trsm":
            assert grid.owner(*n.reads[1]) == plan.exec_rank[n.nid]
    # only panel tiles (matrix O) ever cross ranks
    for lst in list(plan.after_node.values()) + list(plan.before_node.values()):
        for key, m, idx, src, dst in lst:
            assert m is O and src != dst
    assert not plan.before_node                                         # inputs are read where they live
    scripts = [None] * world
    dist.all_gather_object(scripts, plan.describe(rank))
    for a in range(world):
        for b in range(world):
            if a == b:
                continue
            sends = [k for (op, k, peer) in scripts[a] if op == "send" and peer == b]
            recvs = [k for (op, k, peer) in scripts[b] if op == "recv" and peer == a]
            assert sends == recvs, (a, b)
    total = sum(1 for s in scripts for (op, _, _) in s if op == "send")
    assert total == plan.num_transfers
    # a panel tile is sent at most once to each other rank
    fan = {}
    for s in scripts:
        for (op, k, peer) in s:
            if op == "send":
                fan[k] = fan.get(k, 0) + 1
    assert max(fan.values()) <= world - 1

    # inbox slots of the peer-memory exchange: same table on every rank, dense and collision-free per destination
    slot, slot_elems, max_slots = plan.assign_inbox_slots()
    tables = [None] * world
    dist.all_gather_object(tables, sorted((repr(k), v) for k, v in slot.items()))
    assert all(t == tables[0] for t in tables)
    assert slot_elems >= 16 and slot_elems % 16 == 0 and len(slot) == plan.num_transfers
    for d in range(world):
        mine_slots = sorted(v for (k, dst), v in slot.items() if dst == d)
        assert mine_slots == list(range(len(mine_slots))) and len(mine_slots) <= max_slots

    # ---- QR / BDFAC: nodes with two outputs (the tree's trailing updates) must find both output tiles on one rank,
    #      and their transfer plans must pair up like the Cholesky's
    from numpywren_b200 import alg_wrappers
    for which in ("qr", "bdfac"):
        Xq = BigMatrix("dist_%s_in" % which, shape=(24, 24), shard_sizes=(4, 4), device="cpu")
        program, meta = getattr(alg_wrappers, which)(Xq)
        cp = program.program
        qplan = parallel.TransferPlan(cp, grid)
        multi = 0
        for n in cp.nodes:
            owners = {grid.owner(m, idx) for m, idx in n.writes}
            assert len(owners) == 1, (which, n, owners)
            multi += len(n.writes) > 1
        assert multi > 0
        counts = [0] * world
        for r in qplan.exec_rank:
            counts[r] += 1
        assert min(counts) > 0                                            # every rank gets work
        qs = [None] * world
        dist.all_gather_object(qs, qplan.describe(rank))
        for a in range(world):
            for b in range(world):
                if a != b:
                    assert [k for (op, k, peer) in qs[a] if op == "send" and peer == b] == \
                        [k for (op, k, peer) in qs[b] if op == "recv" and peer == a], (which, a, b)
        # inbox slots can be sized: every transferred tile has a declared shape
        slot, slot_elems, max_slots = qplan.assign_inbox_slots()
        assert slot_elems >= 16 and len(slot) == qplan.num_transfers
        for lst in list(qplan.after_node.values()) + list(qplan.before_node.values()):
            for key, m, idx, src, dst in lst:
                assert int(np.prod(parallel._tile_shape(m, idx))) in (16, 32), (which, m.key, idx)

    # ---- binops.gemm across ranks = algs.GEMM_ACC: C tiles never move, only A / B (program inputs) are shipped,
    #      each at most once per destination, and only along its process row (A) or process column (B)
    Ag = BigMatrix("dist_gA", shape=(24, 24), shard_sizes=(4, 4), device="cpu")
    Bg = BigMatrix("dist_gB", shape=(24, 24), shard_sizes=(4, 4), device="cpu")
    program, meta = alg_wrappers.gemm_kloop(Ag, Bg)
    gp = parallel.TransferPlan(program.program, grid)
    assert not gp.after_node                                               # no produced tile crosses ranks
    seen = set()
    for lst in gp.before_node.values():
        for key, m, idx, src, dst in lst:
            assert m is Ag or m is Bg
            assert (key, dst) not in seen
            seen.add((key, dst))
    for n in program.program.nodes:
        if n.call.compute_name == "gemm_acc":
            assert grid.owner(*n.reads[0]) == gp.exec_rank[n.nid] == grid.owner(*n.writes[0])
    gs = [None] * world
    dist.all_gather_object(gs, gp.describe(rank))
    for a in range(world):
        for b in range(world):
            if a != b:
                assert [k for (op, k, peer) in gs[a] if op == "send" and peer == b] == \
                    [k for (op, k, peer) in gs[b] if op == "recv" and peer == a], ("gemm", a, b)

    # ---- failure agreement helper
    assert parallel.allreduce_max_int(rank * 3, torch.device("cpu")) == (world - 1) * 3
    dist.barrier()
    if rank == 0:
        print("DIST_OK", world, total)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
