#!/bin/bash
# Round 2, call 36 (one GPU): compute-sanitizer memcheck + racecheck over every kernel incl. this round's new ones
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 100 python tools/sanitize_small.py 2>&1 | tail -2
( timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py 2>&1 | tail -8; echo "--- racecheck"; timeout 240 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_small.py 2>&1 | tail -12 ) | tee $O/compute_sanitizer.log
