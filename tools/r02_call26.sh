#!/bin/bash
# Round 2, call 26 (one GPU): TSQR workload with the register-resident vs the shared-memory panel kernel (does leaving
# registers free let the update kernels of OTHER leaves run under a panel kernel?)
set -x
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/tsqr_variants.log
for v in 0 1; do
NPW_B200_QR_NO_REG=$v timeout 300 python bench.py --workload tsqr --steps 2 --warmup 1 2>&1 >/dev/null | grep -E "warmup|timed" | sed "s/^/no_reg=$v /" | tee -a $O/tsqr_variants.log
NPW_B200_QR_NO_REG=$v timeout 300 python bench.py --workload tsqr --steps 2 --warmup 1 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('no_reg=$v', d['ms_per_step'], d['roofline']['avg_launch_ms'])" | tee -a $O/tsqr_variants.log
done
