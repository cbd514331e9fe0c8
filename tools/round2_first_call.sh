#!/bin/bash
# FIRST gpurun call of round 2 (one GPU, ~12 min): everything round 1 wrote after its GPU budget ran out gets its
# first run on a B200 here, ordered so that one failure does not hide the rest.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
set -x
mkdir -p gpurun_out
# 1. the parity suite as the driver runs it (QR/BDFAC GPU twins are in tests/test_qr_programs_gpu.py, last in the order)
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
# 2. per-file, so that an early failure in one file cannot mask the others
timeout 600 python -m pytest tests/test_qr_programs_gpu.py -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_qr_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee gpurun_out/smoke.log
# 3. headline bench (unchanged kernels; regression check)
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
# 4. go/no-go for the int8-tensor-core fp64 emulation: library int8 GEMM lower bound
for s in 5 6 8; do timeout 300 python tools/ozaki_lib_probe.py --size 4096 --digits $s 2>&1 | tail -1 | tee -a gpurun_out/ozaki_lib_probe.jsonl; done
# 4a. tcgen05 building blocks in isolation (hand-swizzled shared memory, one MMA group, tcgen05.ld): tells descriptor
#     errors apart from TMA / barrier errors before the product kernel is tried
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/tcgen05_i8_probe tools/tcgen05_i8_probe.cu
timeout 30 /tmp/tcgen05_i8_probe 2>&1 | tail -12 | tee gpurun_out/tcgen05_i8_probe.log
if ! grep -q "PROBE OK" gpurun_out/tcgen05_i8_probe.log; then
  # alternative descriptor encodings (lbo sbo layout version), one process each
  for v in "0 64 2 1" "64 64 2 1" "1 64 2 0" "1 8 2 1" "1 64 1 1"; do
    timeout 30 /tmp/tcgen05_i8_probe $v 2>&1 | tail -4 | tee -a gpurun_out/tcgen05_i8_probe.log
  done
fi
# 4b. the never-run tcgen05 int8 kernel, smallest case first, each under its own timeout (a hang must not cost the box)
for t in "test_split_i8_matches_the_prototype_bit_for_bit" "test_syrk_i8emu_matches_the_prototype[128-64-128-1]" \
         "test_syrk_i8emu_matches_the_prototype[128-64-128-6]" "test_syrk_i8emu_matches_the_prototype" \
         "test_syrk_i8emu_in_place_and_lower_only"; do
  NPW_B200_EXPERIMENTAL=1 timeout 120 python -m pytest "tests/test_i8emu_experimental.py::$t" -m gpu_experimental -x -q 2>&1 | tail -6 | tee -a gpurun_out/i8emu_experimental.log
done
# 5. launch lists of the programs that have never been profiled: QR program, streaming kernels via the GEMM program
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_qr.csv \
    python tools/qr_program_timing.py 4096 512 > gpurun_out/qr_program_under_ncu.log 2>&1
timeout 300 python tools/qr_program_timing.py 8192 1024 2>&1 | tail -3 | tee gpurun_out/qr_program_timing.log
ls -la gpurun_out
