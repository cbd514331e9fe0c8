#!/bin/bash
# Round 2, call 12 (TWO GPUs): TSQR on 2 GPUs with the new leaf kernels + node timeline
set -x
mkdir -p gpurun_out
O=gpurun_out
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $R --nproc-per-node 2 --master-port 29541 bench.py --gpus 2 --workload tsqr --steps 2 --warmup 1 --trace > $O/bench_tsqr_gpus2b.json 2> $O/bench_tsqr_gpus2b.err
tail -4 $O/bench_tsqr_gpus2b.err; grep '^{' $O/bench_tsqr_gpus2b.json | cut -c1-400
