#!/bin/bash
# Round 2, call 6 (one GPU): QR leaf v2 (register-resident panel kernel, stripe-persistent rank update, 4-group T kernel)
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 120 python tools/qr_debug.py 2>&1 | tail -12 | tee $O/qr_debug.log
NPW_B200_QR_NO_REG=1 timeout 120 python tools/qr_debug.py 2>&1 | tail -12 | tee $O/qr_debug_noreg.log
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_qr_programs_gpu.py -q -m gpu -x 2>&1 | tail -15 | tee $O/pytest_qr.log
rm -f $O/qr_leaf_timing.log
for g in 148 128; do NPW_B200_QR_GRID=$g timeout 100 python tools/qr_leaf.py 65536 512 4 2>&1 | tail -2 | sed "s/^/grid $g: /" | tee -a $O/qr_leaf_timing.log; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_qr_leaf.csv python tools/qr_leaf.py 65536 512 1 > /dev/null 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:qr_panel_reg -s 2 -c 1 -o $O/ncu_qr_panel_reg python tools/qr_leaf.py 65536 512 1 > $O/ncu_qr_panel_reg.log 2>&1
timeout 300 $NCU -k regex:rank_update -s 1 -c 1 -o $O/ncu_rank_update python tools/qr_leaf.py 65536 512 1 > $O/ncu_rank_update.log 2>&1
timeout 300 $NCU -k regex:tn_dmma -s 1 -c 1 -o $O/ncu_tn_dmma python tools/qr_leaf.py 65536 512 1 > $O/ncu_tn_dmma.log 2>&1
ls -la $O
