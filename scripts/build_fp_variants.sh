#!/usr/bin/env bash
# Variants of the fingerprint kernel's CTA shape -> warpdemux_b200/lib/var/libwdxfp_<threads>_<minctas>.so (experiments only)
set -eu
cd "$(dirname "$0")/.."
python -m warpdemux_b200.build > /dev/null
L=warpdemux_b200/lib
mkdir -p $L/var
rm -f $L/var/*.so $L/var/*.o
for v in "512 2 16000" "256 4 16000" "256 3 16000" "128 8 8000" "128 6 8000"; do
  set -- $v
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -ccbin /usr/bin/g++ \
    -DWDX_FP_THREADS=$1 -DWDX_FP_MIN_CTAS=$2 -DWDX_FP_MAX_LEN=$3 -Xptxas -v -c -o $L/var/fp_$1_$2.o warpdemux_b200/csrc/wdx_fp.cu 2>&1 | grep -A2 "fingerprint_kernelILb0" | grep -E "spill|Used" | tr '\n' ' '
  echo " <- $1 threads, min $2 CTAs"
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -o $L/var/libwdxfp_$1_$2.so $L/obj/wdx_b200.o $L/var/fp_$1_$2.o $L/obj/wdx_cnn.o $L/obj/wdx_validate.o
done
