#!/usr/bin/env bash
# phase-clock build of the fingerprint kernel -> warpdemux_b200/lib/var/libwdxfp_prof.so (experiments only)
set -eu
cd "$(dirname "$0")/.."
python -m warpdemux_b200.build > /dev/null
L=warpdemux_b200/lib
mkdir -p $L/var
rm -f $L/var/*.so $L/var/*.o
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -ccbin /usr/bin/g++ \
    -DWDX_FP_PROF ${WDX_FP_EXTRA:-} -c -o $L/var/fp_prof.o warpdemux_b200/csrc/wdx_fp.cu
nvcc -shared -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -o $L/var/libwdxfp_prof.so $L/obj/wdx_b200.o $L/var/fp_prof.o $L/obj/wdx_cnn.o $L/obj/wdx_validate.o
