#!/bin/bash
# Round 2, call 2 (one GPU): new bench (N=131072 on one GPU, parity_vs_oracle, e2e ring, composed CPU baseline), reference
# arm, int8-emulated syrk timings.
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee $O/pytest_gpu.log
timeout 900 python bench.py --steps 2 --warmup 1 > $O/bench.json 2> $O/bench.err; tail -5 $O/bench.err; cut -c1-1500 $O/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; tail -3 $O/bench_ref.err; cut -c1-1200 $O/bench_ref.json
for s in 6 7 8; do NPW_B200_EXPERIMENTAL=1 timeout 120 python tools/syrk_i8emu_timing.py $s 4096 2>&1 | tail -1 | tee -a $O/syrk_i8emu_timing.jsonl; done
NPW_B200_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_i8emu_experimental.py -m gpu_experimental -q 2>&1 | tail -4 | tee $O/i8emu_experimental.log
NPW_B200_SYRK=i8emu NPW_B200_I8_DIGITS=8 NPW_B200_BENCH_NO_E2E=1 timeout 400 python bench.py --size 65536 --steps 2 --warmup 1 > $O/bench_i8emu8.json 2> $O/bench_i8emu8.err; tail -5 $O/bench_i8emu8.err; cut -c1-1500 $O/bench_i8emu8.json
ls -la $O
