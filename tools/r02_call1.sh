#!/bin/bash
# Round 2, call 1 (one GPU): regression of the round-1 state + per-kernel ncu captures + probes.
set -x
mkdir -p gpurun_out
O=gpurun_out
( nproc; free -g; nvidia-smi -L; python -c "import os;print(os.cpu_count())" ) > $O/box.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee $O/pytest_gpu.log
timeout 400 python bench.py --steps 2 --warmup 3 > $O/bench.json 2> $O/bench.err; tail -3 $O/bench.err; cut -c1-900 $O/bench.json
# N=131072 on ONE GPU: does it fit and what is the step time
NPW_B200_BENCH_NO_E2E=1 timeout 400 python bench.py --size 131072 --steps 1 --warmup 1 --no-cpu > $O/bench_n131072_1gpu.json 2> $O/bench_n131072_1gpu.err; tail -3 $O/bench_n131072_1gpu.err; cut -c1-600 $O/bench_n131072_1gpu.json
# timings
timeout 200 python tools/factor_timing.py 2>&1 | tee $O/factor_timing.log
timeout 200 python tools/qr_leaf.py 65536 512 3 2>&1 | tee $O/qr_leaf_timing.log
# per-kernel ncu --set full
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:potf2_inv -c 3 -o $O/ncu_potf2_inv python tools/potrf_only.py 4096 > $O/ncu_potf2.log 2>&1
timeout 300 $NCU -k regex:qr_panel -s 2 -c 2 -o $O/ncu_qr_panel python tools/qr_leaf.py 65536 512 1 > $O/ncu_qr_panel.log 2>&1
timeout 300 $NCU -k regex:tn_partial -s 2 -c 2 -o $O/ncu_tn_partial python tools/qr_leaf.py 65536 512 1 > $O/ncu_tn_partial.log 2>&1
timeout 300 $NCU -s 5 -c 12 -o $O/ncu_streaming python tools/streaming_only.py 4096 > $O/ncu_streaming.log 2>&1
# launch lists: potrf + trsm tile ops, QR leaf
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_potrf.csv python tools/potrf_only.py 4096 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_qr_leaf.csv python tools/qr_leaf.py 65536 512 1 > /dev/null 2>&1
# int8 emulation: call 1 of at most 3
for s in 6 8; do timeout 200 python tools/ozaki_lib_probe.py --size 4096 --digits $s 2>&1 | tail -1 | tee -a $O/ozaki_lib_probe.jsonl; done
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/tcgen05_i8_probe tools/tcgen05_i8_probe.cu
timeout 30 /tmp/tcgen05_i8_probe 2>&1 | tail -12 | tee $O/tcgen05_i8_probe.log
if ! grep -q "PROBE OK" $O/tcgen05_i8_probe.log; then
  for v in "0 64 2 1" "64 64 2 1" "1 64 2 0" "1 8 2 1" "1 64 1 1"; do
    timeout 30 /tmp/tcgen05_i8_probe $v 2>&1 | tail -4 | tee -a $O/tcgen05_i8_probe.log
  done
fi
for t in "test_split_i8_matches_the_prototype_bit_for_bit" "test_syrk_i8emu_matches_the_prototype[128-64-128-1]" \
         "test_syrk_i8emu_matches_the_prototype[128-64-128-6]" "test_syrk_i8emu_matches_the_prototype"; do
  NPW_B200_EXPERIMENTAL=1 timeout 90 python -m pytest "tests/test_i8emu_experimental.py::$t" -m gpu_experimental -x -q 2>&1 | tail -6 | tee -a $O/i8emu_experimental.log
done
ls -la $O
