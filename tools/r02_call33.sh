#!/bin/bash
# Round 2, call 33 (one GPU): short row loops in the panel kernel for CTAs with <= 128 rows (TSQR merges)
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 120 python tools/qr_debug.py 2>&1 | tail -12 | tee $O/qr_debug.log
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_qr_programs_gpu.py tests/test_algs_gpu.py -q -m gpu -x 2>&1 | tail -3 | tee $O/pytest_qr.log
timeout 100 python tools/qr_leaf.py 65536 512 4 2>&1 | tail -1 | tee $O/qr_leaf_timing.log
timeout 100 python tools/qr_leaf.py 1024 512 6 2>&1 | tail -1 | tee -a $O/qr_leaf_timing.log
timeout 300 python bench.py --workload tsqr --steps 2 --warmup 1 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('tsqr 1 gpu', d['ms_per_step'], d['config']['parity_vs_golden']['rel_fro_R'])" | tee -a $O/qr_leaf_timing.log
