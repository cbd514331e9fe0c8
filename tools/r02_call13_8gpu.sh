#!/bin/bash
# Round 2, call 13 (EIGHT GPUs, charged 8x — keep it short): multi-GPU parity tests at 2/4/8 ranks, BASELINE configs 3, 4, 5
set -x
mkdir -p gpurun_out
O=gpurun_out
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L > $O/box8.txt
timeout 400 python -m pytest tests/test_multi_gpu.py -q -m gpu 2>&1 | tail -8 | tee $O/pytest_multi_gpu_8.log
timeout 200 $R --nproc-per-node 8 --master-port 29551 bench.py --gpus 8 --workload tsqr --steps 3 --warmup 2 --trace > $O/bench_tsqr_gpus8.json 2> $O/bench_tsqr_gpus8.err
tail -3 $O/bench_tsqr_gpus8.err; grep '^{' $O/bench_tsqr_gpus8.json | cut -c1-500
timeout 300 $R --nproc-per-node 8 --master-port 29552 bench.py --gpus 8 --workload gemm --steps 1 --warmup 1 > $O/bench_gemm_gpus8.json 2> $O/bench_gemm_gpus8.err
tail -3 $O/bench_gemm_gpus8.err; grep '^{' $O/bench_gemm_gpus8.json | cut -c1-700
timeout 400 $R --nproc-per-node 8 --master-port 29553 bench.py --gpus 8 --steps 2 --warmup 1 --trace > $O/bench_gpus8.json 2> $O/bench_gpus8.err
tail -6 $O/bench_gpus8.err; grep '^{' $O/bench_gpus8.json | cut -c1-3000
ls -la $O
