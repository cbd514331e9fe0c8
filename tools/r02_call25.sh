#!/bin/bash
# Round 2, call 25 (one GPU): phase breakdown of the panel kernel from a -DNPW_QR_PROFILE build (built on the box; the
# product library in the tree is not touched: the profile objects go to /tmp and the .so is swapped only for this process)
set -x
mkdir -p gpurun_out
O=gpurun_out
cd numpywren_b200/csrc
mkdir -p /tmp/prof
for f in npw_common npw_gemm_f64 npw_elementwise npw_factor_f64 npw_ozaki_i8; do cp build/$f.o /tmp/prof/; done
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DNPW_QR_PROFILE -c npw_qr_f64.cu -o /tmp/prof/npw_qr_f64.o
cp ../lib/libnpw_b200.so /tmp/prof/libnpw_b200.so.orig
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/libnpw_b200.so /tmp/prof/*.o -cudart shared
cd ../..
timeout 120 python tools/qr_panel_profile.py 2>&1 | tail -8 | tee $O/qr_panel_profile.log
cp /tmp/prof/libnpw_b200.so.orig numpywren_b200/lib/libnpw_b200.so
