#!/usr/bin/env bash
# Variants of the validation kernel's CTA shape -> warpdemux_b200/lib/var/libwdx_<threads>_<minctas>.so (experiments only)
set -eu
cd "$(dirname "$0")/.."
python -m warpdemux_b200.build > /dev/null
L=warpdemux_b200/lib
mkdir -p $L/var
for v in "512 2" "512 1" "256 3" "256 4" "256 2" "128 4"; do
  set -- $v
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -ccbin /usr/bin/g++ \
    -DWDX_VAL_THREADS=$1 -DWDX_VAL_MIN_CTAS=$2 -Xptxas -v -c -o $L/var/val_$1_$2.o warpdemux_b200/csrc/wdx_validate.cu 2>&1 | grep -A2 "validate_kernelENS_7ValArgs" | grep -E "spill|Used" | tr '\n' ' '
  echo " <- $1 threads, min $2 CTAs"
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -o $L/var/libwdx_$1_$2.so $L/obj/wdx_b200.o $L/obj/wdx_fp.o $L/obj/wdx_cnn.o $L/var/val_$1_$2.o
done
