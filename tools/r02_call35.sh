#!/bin/bash
# Round 2, call 35 (one GPU): final state — full -m gpu suite, smoke, ncu --set full of the final QR kernels, leaf launch list
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee $O/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke.log
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 200 $NCU -k regex:qr_panel_reg -s 2 -c 1 -o $O/ncu_qr_panel_reg_final python tools/qr_leaf.py 65536 512 1 > /dev/null 2>&1
timeout 200 $NCU -k regex:rank_update -s 1 -c 1 -o $O/ncu_rank_update_final python tools/qr_leaf.py 65536 512 1 > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_qr_leaf.csv python tools/qr_leaf.py 65536 512 1 > /dev/null 2>&1
timeout 100 python tools/qr_leaf.py 65536 512 4 2>&1 | tail -1 | tee $O/qr_leaf_timing.log
