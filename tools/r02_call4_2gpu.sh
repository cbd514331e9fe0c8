#!/bin/bash
# Round 2, call 4 (TWO GPUs, charged 2x): first hardware run of the multi-GPU tests of this round and of the tsqr / gemm
# bench workloads on more than one GPU; Cholesky N=131072 at 2 GPUs with e2e + utilisation trace.
set -x
mkdir -p gpurun_out
O=gpurun_out
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L | tee $O/box2.txt
timeout 600 python -m pytest tests/test_multi_gpu.py -q -m gpu 2>&1 | tail -15 | tee $O/pytest_multi_gpu_2.log
timeout 300 $R --nproc-per-node 2 --master-port 29541 bench.py --gpus 2 --workload tsqr --steps 2 --warmup 1 > $O/bench_tsqr_gpus2.json 2> $O/bench_tsqr_gpus2.err
tail -5 $O/bench_tsqr_gpus2.err; cut -c1-1500 $O/bench_tsqr_gpus2.json
timeout 300 $R --nproc-per-node 2 --master-port 29542 bench.py --gpus 2 --workload gemm --steps 1 --warmup 1 > $O/bench_gemm_gpus2.json 2> $O/bench_gemm_gpus2.err
tail -5 $O/bench_gemm_gpus2.err; cut -c1-1500 $O/bench_gemm_gpus2.json
timeout 600 $R --nproc-per-node 2 --master-port 29543 bench.py --gpus 2 --steps 2 --warmup 1 --trace > $O/bench_gpus2.json 2> $O/bench_gpus2.err
tail -8 $O/bench_gpus2.err; cut -c1-3000 $O/bench_gpus2.json
timeout 200 $R --nproc-per-node 2 --master-port 29544 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > $O/bench_ref_gpus2.json 2> $O/bench_ref_gpus2.err
tail -3 $O/bench_ref_gpus2.err; cut -c1-600 $O/bench_ref_gpus2.json
ls -la $O
