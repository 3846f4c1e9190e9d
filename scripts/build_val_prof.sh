#!/usr/bin/env bash
# phase-clock build of the validation kernel -> warpdemux_b200/lib/var/libwdxval_prof.so (experiments only)
set -eu
cd "$(dirname "$0")/.."
python -m warpdemux_b200.build > /dev/null
L=warpdemux_b200/lib
mkdir -p $L/var
rm -f $L/var/*.so $L/var/*.o
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -ccbin /usr/bin/g++ \
    -DWDX_FP_PROF -c -o $L/var/val_prof.o warpdemux_b200/csrc/wdx_validate.cu
nvcc -shared -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -o $L/var/libwdxval_prof.so $L/obj/wdx_b200.o $L/obj/wdx_fp.o $L/obj/wdx_cnn.o $L/var/val_prof.o
