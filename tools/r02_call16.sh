#!/bin/bash
# Round 2, call 16 (one GPU): full -m gpu suite (int8 tests promoted, tall-tile QR variants), trsm with fork/join row
# halves, smoke(), short bench at N=65536.
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee $O/pytest_gpu.log
timeout 200 python tools/factor_timing.py 2>&1 | tee $O/factor_timing.log
NPW_B200_TRSM_SPLIT=0 timeout 200 python tools/factor_timing.py 2>&1 | tail -1 | sed 's/^/nosplit: /' | tee -a $O/factor_timing.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke.log
NPW_B200_BENCH_NO_E2E=1 timeout 400 python bench.py --size 65536 --steps 3 --warmup 2 --no-cpu > $O/bench_n65536.json 2> $O/bench_n65536.err; tail -3 $O/bench_n65536.err; grep '^{' $O/bench_n65536.json | cut -c1-300
ls -la $O
