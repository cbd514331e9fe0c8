"""Time the tiled QR program (alg_wrappers.qr, Householder semantics) on one GPU: python tools/qr_program_timing.py N TILE."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from numpywren_b200 import alg_wrappers, job_runner, kernels, qr  # noqa: E402
from numpywren_b200 import lambdapack as lp  # noqa: E402
from numpywren_b200.matrix import BigMatrix  # noqa: E402


def main():
    n, b = int(sys.argv[1]), int(sys.argv[2])
    qr.set_qr_semantics("householder")
    torch.cuda.set_device(0)
    A = BigMatrix("qrt_A", shape=(n, n), shard_sizes=(b, b))
    A.free()
    for bi in A.block_idxs:
        t = torch.empty(A.block_shape(*bi), dtype=torch.float64, device="cuda")
        kernels.fill_random(t, 7, bi[0] * b, bi[1] * b)
        A._put_block_ref(t, *bi)
    for rep in range(2):
        program, meta = alg_wrappers.qr(A)
        for m in meta["outputs"] + meta["intermediates"]:
            m.free()
        torch.cuda.synchronize()
        t0 = time.time()
        program.start()
        job_runner.lambdapack_run(program, timeout=600, free_intermediates=True)
        torch.cuda.synchronize()
        dt = time.time() - t0
        assert program.program_status() == lp.PS.SUCCESS
        flops = 4.0 * n ** 3 / 3.0
        print(f"qr N={n} tile={b}: {len(program.program.nodes)} tile tasks, {dt * 1e3:.1f} ms, {flops / dt / 1e12:.2f} TFLOP/s (4N^3/3)")
    Rs = meta["outputs"][0]
    R00 = Rs.get_block(0, 0, 0).cpu().numpy()
    print("R[0,0] diag head", np.diag(R00)[:4])


if __name__ == "__main__":
    main()
