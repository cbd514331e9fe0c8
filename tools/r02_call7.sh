#!/bin/bash
# Round 2, call 7 (one GPU): QR leaf v3 (arithmetic lane roles in the register-resident panel kernel)
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 120 python tools/qr_debug.py 2>&1 | tail -12 | tee $O/qr_debug.log
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_qr_programs_gpu.py -q -m gpu -x 2>&1 | tail -15 | tee $O/pytest_qr.log
rm -f $O/qr_leaf_timing.log
timeout 100 python tools/qr_leaf.py 65536 512 4 2>&1 | tail -2 | tee -a $O/qr_leaf_timing.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_qr_leaf.csv python tools/qr_leaf.py 65536 512 1 > /dev/null 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:qr_panel_reg -s 2 -c 1 -o $O/ncu_qr_panel_reg python tools/qr_leaf.py 65536 512 1 > $O/ncu_qr_panel_reg.log 2>&1
ls -la $O
