#!/bin/bash
# Round 2, call 22 (one GPU): L2 eviction hints in the GEMM kernel (A panel evict_last, C evict_first) x raster group:
# DRAM bytes per 4096^3 syrk from ncu, kernel time, parity of the kernel tests with the hints on
set -x
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/syrk_l2hint.txt
for h in 0 1; do for r in 8 16; do
NPW_B200_L2HINT=$h NPW_B200_RASTER=$r timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:gemm_nt_tma -s 4 -c 2 --csv --log-file $O/syrk_h${h}_r$r.csv python tools/syrk_only.py 4096 > /dev/null 2>&1
echo "hint=$h raster=$r" >> $O/syrk_l2hint.txt
grep -E "dram__bytes|time_duration|hit_rate" $O/syrk_h${h}_r$r.csv | awk -F'","' '{print "   ", $(NF-2), $(NF-1), $NF}' | tail -4 >> $O/syrk_l2hint.txt
done; done
cat $O/syrk_l2hint.txt
NPW_B200_L2HINT=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x 2>&1 | tail -2 | tee $O/pytest_l2hint.log
for h in 0 1; do NPW_B200_L2HINT=$h NPW_B200_BENCH_NO_E2E=1 timeout 300 python bench.py --size 65536 --steps 3 --warmup 2 --no-cpu 2>&1 >/dev/null | grep timed | sed "s/^/l2hint=$h /" | tee -a $O/syrk_l2hint.txt; done
