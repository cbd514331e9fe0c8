#!/bin/bash
# Round 2, call 38 (one GPU): the complete -m gpu suite in its final composition
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
