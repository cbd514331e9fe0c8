#!/bin/bash
# Round 2, call 10 (one GPU): QR leaf v6 (gpu-scope packets, no error-word read on the fast path) + the TSQR workload on one GPU
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 120 python tools/qr_debug.py 2>&1 | tail -12 | tee $O/qr_debug.log
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_qr_programs_gpu.py -q -m gpu -x 2>&1 | tail -3 | tee $O/pytest_qr.log
rm -f $O/qr_leaf_timing.log
timeout 100 python tools/qr_leaf.py 65536 512 4 2>&1 | tail -2 | tee -a $O/qr_leaf_timing.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_qr_leaf.csv python tools/qr_leaf.py 65536 512 1 > /dev/null 2>&1
timeout 300 python bench.py --workload tsqr --steps 2 --warmup 1 > $O/bench_tsqr_gpus1.json 2> $O/bench_tsqr_gpus1.err; tail -3 $O/bench_tsqr_gpus1.err; cut -c1-400 $O/bench_tsqr_gpus1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file $O/launches_tsqr.csv python bench.py --workload tsqr --size 1048576 --steps 1 --warmup 1 > /dev/null 2>&1
ls -la $O
