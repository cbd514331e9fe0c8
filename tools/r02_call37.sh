#!/bin/bash
# Round 2, call 37 (one GPU): the INTEGRATION §2 reference-side stub (ctypes + cudart, no torch) against the reference's fixture
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_reference_stub_gpu.py -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_stub.log
