#!/bin/bash
# Round 2, call 21 (one GPU): single-hop gather for small grids, R rows written by the rank update
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 120 python tools/qr_debug.py 2>&1 | tail -12 | tee $O/qr_debug.log
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_qr_programs_gpu.py tests/test_algs_gpu.py -q -m gpu -x 2>&1 | tail -3 | tee $O/pytest_qr.log
timeout 100 python tools/qr_leaf.py 65536 512 4 2>&1 | tail -2 | tee $O/qr_leaf_timing.log
timeout 100 python tools/qr_leaf.py 1024 512 6 2>&1 | tail -2 | tee -a $O/qr_leaf_timing.log
timeout 300 python bench.py --workload tsqr --steps 2 --warmup 1 > $O/bench_tsqr_gpus1.json 2> $O/bench_tsqr_gpus1.err; tail -2 $O/bench_tsqr_gpus1.err; grep '^{' $O/bench_tsqr_gpus1.json | cut -c1-300
