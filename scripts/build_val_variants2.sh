#!/usr/bin/env bash
# Variants of the validation kernel (CTA shape x row in shared memory / left in global memory) -> warpdemux_b200/lib/var/
set -eu
cd "$(dirname "$0")/.."
python -m warpdemux_b200.build > /dev/null
L=warpdemux_b200/lib
rm -rf $L/var; mkdir -p $L/var
for v in "512 2 0" "128 8 1" "128 6 1" "256 4 1" "256 3 1" "512 2 1" "64 16 1"; do
  set -- $v
  G=""; [ "$3" = "1" ] && G="-DWDX_VAL_GLOBAL_ROW"
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -ccbin /usr/bin/g++ \
    -DWDX_VAL_THREADS=$1 -DWDX_VAL_MIN_CTAS=$2 $G -Xptxas -v -c -o $L/var/val_$1_$2_$3.o warpdemux_b200/csrc/wdx_validate.cu 2>&1 | grep -A2 "validate_kernelENS_7ValArgs" | grep -E "spill|Used" | tr '\n' ' '
  echo " <- $1 threads, min $2 CTAs, global row $3"
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -o $L/var/libwdx_$1_$2_$3.so $L/obj/wdx_b200.o $L/obj/wdx_fp.o $L/obj/wdx_cnn.o $L/var/val_$1_$2_$3.o
done
