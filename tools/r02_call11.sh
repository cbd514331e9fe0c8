#!/bin/bash
# Round 2, call 11 (one GPU): int8 syrk with raster + 6 X slots + interleaved digit order; golden parity in the tsqr / gemm
# bench arms; syrk DRAM traffic with raster groups 8 / 16 (ncu dram bytes).
set -x
mkdir -p gpurun_out
O=gpurun_out
NPW_B200_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_i8emu_experimental.py -m gpu_experimental -q 2>&1 | tail -4 | tee $O/i8emu_experimental.log
rm -f $O/syrk_i8emu_timing3.jsonl
for s in 6 7 8; do NPW_B200_EXPERIMENTAL=1 timeout 120 python tools/syrk_i8emu_timing.py $s 4096 2>&1 | tail -1 | tee -a $O/syrk_i8emu_timing3.jsonl; done
timeout 300 python bench.py --workload tsqr --size 1048576 --steps 1 --warmup 1 > $O/bench_tsqr_small.json 2> $O/bench_tsqr_small.err; tail -3 $O/bench_tsqr_small.err; cut -c1-900 $O/bench_tsqr_small.json
timeout 300 python bench.py --workload gemm --size 32768 --steps 1 --warmup 1 > $O/bench_gemm_small.json 2> $O/bench_gemm_small.err; tail -3 $O/bench_gemm_small.err; cut -c1-900 $O/bench_gemm_small.json
for r in 8 16; do
NPW_B200_RASTER=$r timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:gemm_nt_tma -s 2 -c 2 --csv --log-file $O/syrk_raster$r.csv python tools/syrk_only.py 4096 > /dev/null 2>&1
grep -E "dram__bytes|time_duration|hit_rate" $O/syrk_raster$r.csv | cut -d, -f 5,13- | tail -8
done
NCU="ncu --set full --clock-control none --import-source on -f"
NPW_B200_EXPERIMENTAL=1 timeout 300 $NCU -k regex:ozaki_syrk -s 2 -c 1 -o $O/ncu_ozaki_syrk2 python tools/syrk_i8emu_timing.py 7 4096 > $O/ncu_ozaki2.log 2>&1
ls -la $O
