#!/bin/bash
# Round 2, call 30 (one GPU): panel kernel with two rows per warp instruction (PAIR) vs one
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 120 python tools/qr_debug.py 2>&1 | tail -12 | tee $O/qr_debug.log
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_qr_programs_gpu.py tests/test_algs_gpu.py -q -m gpu -x 2>&1 | tail -3 | tee $O/pytest_qr.log
rm -f $O/qr_leaf_timing.log
for v in 1 0; do NPW_B200_QR_PAIR=$v timeout 100 python tools/qr_leaf.py 65536 512 4 2>&1 | tail -1 | sed "s/^/pair=$v /" | tee -a $O/qr_leaf_timing.log; done
timeout 100 python tools/qr_leaf.py 1024 512 6 2>&1 | tail -1 | tee -a $O/qr_leaf_timing.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_qr_leaf.csv python tools/qr_leaf.py 65536 512 1 > /dev/null 2>&1
