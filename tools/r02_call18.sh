#!/bin/bash
# Round 2, call 18 (one GPU): 64-row vs 128-row chunks in the potrf / trsm panel kernel (SM time stolen from the syrks
# vs chain latency): tile timings and the whole Cholesky at N=65536
set -x
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/panel_tr_ab.log
for v in 12288 0; do
NPW_B200_PANEL_SMALL_ROWS=$v timeout 200 python tools/factor_timing.py 2>&1 | tail -1 | sed "s/^/small_rows=$v /" | tee -a $O/panel_tr_ab.log
NPW_B200_PANEL_SMALL_ROWS=$v NPW_B200_BENCH_NO_E2E=1 timeout 300 python bench.py --size 65536 --steps 3 --warmup 2 --no-cpu 2>&1 >/dev/null | grep timed | sed "s/^/small_rows=$v /" | tee -a $O/panel_tr_ab.log
done
