#!/bin/bash
# Round 2, call 3 (one GPU): potrf/trsm rewrite (panel_kernel) regression + timings; int8-emulated syrk timings with a
# proper warm-up and one ncu --set full capture of it.
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | tee $O/pytest_gpu.log
timeout 200 python tools/factor_timing.py 2>&1 | tee $O/factor_timing.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_potrf.csv python tools/potrf_only.py 4096 > /dev/null 2>&1
for s in 6 7 8; do NPW_B200_EXPERIMENTAL=1 timeout 120 python tools/syrk_i8emu_timing.py $s 4096 2>&1 | tail -1 | tee -a $O/syrk_i8emu_timing2.jsonl; done
NPW_B200_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_i8emu_experimental.py -m gpu_experimental -q 2>&1 | tail -4 | tee $O/i8emu_experimental.log
NCU="ncu --set full --clock-control none --import-source on -f"
NPW_B200_EXPERIMENTAL=1 timeout 300 $NCU -k regex:ozaki_syrk -s 2 -c 1 -o $O/ncu_ozaki_syrk python tools/syrk_i8emu_timing.py 7 4096 > $O/ncu_ozaki.log 2>&1
timeout 300 $NCU -k regex:panel_kernel -s 40 -c 2 -o $O/ncu_panel_kernel python tools/potrf_only.py 4096 > $O/ncu_panel.log 2>&1
ls -la $O
