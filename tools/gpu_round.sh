#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list + full capture of the dominant kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee gpurun_out/smoke.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; cat gpurun_out/bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --size 16384 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_nt_tma -s 10 -c 3 -f -o gpurun_out/prof_syrk \
    python tools/syrk_only.py 4096 16 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
