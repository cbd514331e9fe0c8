#!/bin/bash
# Round 2, call 28 (one GPU): what the driver runs at round end — full -m gpu suite, smoke(), default bench (fewer steps),
# reference arm
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee $O/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > $O/bench.json 2> $O/bench.err; tail -4 $O/bench.err; grep '^{' $O/bench.json | cut -c1-600
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; grep '^{' $O/bench_ref.json | cut -c1-400
