#!/bin/bash
# SECOND gpurun call of round 2 (8 GPUs, charged 8x: keep it short, ~6 min):
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 900 -- 'bash tools/round2_multi_gpu_call.sh'
set -x
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
# 1. multi-GPU parity incl. the QR / BDFAC / binops.gemm paths written after round 1's GPU budget was spent
timeout 600 python -m pytest tests/test_multi_gpu.py -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_multi_gpu.log
# 2. headline at 8 GPUs: the static plan is now built before the timed region (config.plan_s) — expect ~2.7-2.8 s
timeout 300 $R --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps 2 --warmup 2 > gpurun_out/bench_gpus8.json 2> gpurun_out/bench_gpus8.err
tail -3 gpurun_out/bench_gpus8.err; cut -c1-600 gpurun_out/bench_gpus8.json
# 3. the same with host buffers (opt-in path, first run on hardware)
NPW_B200_BENCH_E2E=1 timeout 400 $R --nproc-per-node 8 --master-port 29532 bench.py --gpus 8 --steps 1 --warmup 1 > gpurun_out/bench_gpus8_e2e.json 2> gpurun_out/bench_gpus8_e2e.err
tail -3 gpurun_out/bench_gpus8_e2e.err; python -c "import json;d=json.load(open('gpurun_out/bench_gpus8_e2e.json'));print(d['value'], d['e2e'])"
# 4. BASELINE config 5 (GEMM N=131072, tile 8192) across 8 GPUs
timeout 400 $R --nproc-per-node 8 --master-port 29533 tools/config5_multi_gpu.py 131072 8192 2>&1 | tail -4 | tee gpurun_out/config5_gpus8.log
ls -la gpurun_out
