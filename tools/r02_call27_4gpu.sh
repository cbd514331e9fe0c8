#!/bin/bash
# Round 2, call 27 (FOUR GPUs): final-state check of the multi-GPU paths that changed since the 8-GPU calls: TSQR
# (panel kernel polling / one-hop gather), multi-GPU tests at 2 and 4 ranks
set -x
mkdir -p gpurun_out
O=gpurun_out
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
NPW_B200_SIGNAL_TIMEOUT_MS=20000 timeout 150 $R --nproc-per-node 4 --master-port 29581 bench.py --gpus 4 --workload tsqr --steps 3 --warmup 2 > $O/bench_tsqr_gpus4b.json 2> $O/bench_tsqr_gpus4b.err
grep -v "^\*\|OMP_NUM" $O/bench_tsqr_gpus4b.err | tail -3 | cut -c1-300; grep '^{' $O/bench_tsqr_gpus4b.json | cut -c1-300
timeout 400 python -m pytest tests/test_multi_gpu.py -q -m gpu 2>&1 | tail -3 | tee $O/pytest_multi_gpu_4.log
