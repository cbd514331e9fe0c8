#!/bin/bash
# Round 2, call 17 (one GPU): A/B of the trsm fork/join at the whole-Cholesky level (N=65536)
set -x
mkdir -p gpurun_out
O=gpurun_out
for v in 1 0 1 0; do
NPW_B200_TRSM_SPLIT=$v NPW_B200_BENCH_NO_E2E=1 timeout 300 python bench.py --size 65536 --steps 3 --warmup 2 --no-cpu 2>&1 >/dev/null | grep timed | sed "s/^/split=$v /" | tee -a $O/trsm_split_ab.log
done
