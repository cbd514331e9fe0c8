#!/bin/bash
# Round 2, call 32 (one GPU): launch breakdown of a TSQR merge (geqrt of the 1024 x 512 stack)
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_qr_merge.csv python tools/qr_leaf.py 1024 512 1 > /dev/null 2>&1
timeout 100 python tools/qr_leaf.py 1024 512 6 2>&1 | tail -2
