#!/bin/bash
# Round 2, call 29 (EIGHT GPUs, short): BASELINE config 4 (TSQR) with the final kernels
set -x
mkdir -p gpurun_out
O=gpurun_out
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
NPW_B200_SIGNAL_TIMEOUT_MS=20000 timeout 100 $R --nproc-per-node 8 --master-port 29591 bench.py --gpus 8 --workload tsqr --steps 5 --warmup 3 --trace > $O/bench_tsqr_gpus8b.json 2> $O/bench_tsqr_gpus8b.err
grep -v "^\*\|OMP_NUM" $O/bench_tsqr_gpus8b.err | tail -3 | cut -c1-300; grep '^{' $O/bench_tsqr_gpus8b.json | cut -c1-300
