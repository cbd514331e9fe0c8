"""BASELINE config 5 on several GPUs: GEMM N x N fp64, tile 8192, legacy binops.gemm schedule (algs.GEMM_ACC on the engine).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
      tools/config5_multi_gpu.py [N=131072] [TILE=8192]
Every rank generates the A and B tiles it owns on its GPU; rank 0 prints time and TFLOP/s (2 N^3 flops)."""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from numpywren_b200 import alg_wrappers, job_runner, kernels, parallel  # noqa: E402
from numpywren_b200 import lambdapack as lp  # noqa: E402
from numpywren_b200.matrix import BigMatrix  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
    b = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    grid = parallel.init_from_env("nccl")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    nb = n // b
    A = BigMatrix("c5_A", shape=(n, n), shard_sizes=(b, b), device=dev)
    B = BigMatrix("c5_B", shape=(n, n), shard_sizes=(b, b), device=dev)
    for i in range(nb):
        for k in range(nb):
            for m, seed in ((A, 1), (B, 2)):
                if grid.is_mine(m, (i, k)):
                    t = torch.empty(b, b, dtype=torch.float64, device=dev)
                    kernels.fill_random(t, seed, i * b, k * b)
                    m._put_block_ref(t, i, k)
    for rep in range(2):
        program, meta = alg_wrappers.gemm_kloop(A, B, out_key=f"c5_C{rep}")
        plan_s = job_runner.prepare(program, streams=8)
        torch.cuda.synchronize()
        dist.barrier(device_ids=[dev.index])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        program.start()
        job_runner.lambdapack_run(program, timeout=3600, streams=8)
        e1.record()
        e1.synchronize()
        assert program.program_status() == lp.PS.SUCCESS
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if grid.rank == 0:
            dt = float(ms.item()) * 1e-3
            print(f"GEMM N={n} tile {b} on {grid.world} GPUs ({grid.P}x{grid.Q}): {len(program.program.nodes)} tile tasks, "
                  f"plan {plan_s:.2f} s, {dt * 1e3:.0f} ms -> {2.0 * n ** 3 / dt * 1e-12:.1f} TFLOP/s", flush=True)
        for mm in meta["outputs"] + meta["intermediates"]:
            mm.free()
    dist.barrier(device_ids=[dev.index])
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
