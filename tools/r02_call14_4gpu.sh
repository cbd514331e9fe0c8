#!/bin/bash
# Round 2, call 14 (FOUR GPUs): TSQR with round-robin merges after the register-headroom fix of the panel kernel
set -x
mkdir -p gpurun_out
O=gpurun_out
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
NPW_B200_SIGNAL_TIMEOUT_MS=15000 timeout 150 $R --nproc-per-node 4 --master-port 29561 bench.py --gpus 4 --workload tsqr --steps 2 --warmup 1 --trace > $O/bench_tsqr_gpus4.json 2> $O/bench_tsqr_gpus4.err
grep -v "^\*\|OMP_NUM" $O/bench_tsqr_gpus4.err | tail -5 | cut -c1-300; grep '^{' $O/bench_tsqr_gpus4.json | cut -c1-400
