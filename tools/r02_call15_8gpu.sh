#!/bin/bash
# Round 2, call 15 (EIGHT GPUs): BASELINE configs 4 (TSQR) and 5 (GEMM N=131072 tile 8192) after the fixes of call 13
set -x
mkdir -p gpurun_out
O=gpurun_out
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
NPW_B200_SIGNAL_TIMEOUT_MS=20000 timeout 120 $R --nproc-per-node 8 --master-port 29571 bench.py --gpus 8 --workload tsqr --steps 3 --warmup 2 --trace > $O/bench_tsqr_gpus8.json 2> $O/bench_tsqr_gpus8.err
grep -v "^\*\|OMP_NUM" $O/bench_tsqr_gpus8.err | tail -4 | cut -c1-300; grep '^{' $O/bench_tsqr_gpus8.json | cut -c1-400
NPW_B200_SIGNAL_TIMEOUT_MS=60000 timeout 240 $R --nproc-per-node 8 --master-port 29572 bench.py --gpus 8 --workload gemm --steps 1 --warmup 1 > $O/bench_gemm_gpus8.json 2> $O/bench_gemm_gpus8.err
grep -v "^\*\|OMP_NUM" $O/bench_gemm_gpus8.err | tail -4 | cut -c1-300; grep '^{' $O/bench_gemm_gpus8.json | cut -c1-700
