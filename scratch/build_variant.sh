#!/bin/bash
# build_variant.sh NAME "EXTRA_NVCC_FLAGS": builds scratch/variants/libsvsb200_NAME.so (experiments; select with SVSB200_LIB)
set -e
cd "$(dirname "$0")/../scalablevectorsearch_b200/csrc"
NAME=$1; EXTRA=$2
OUT=../../scratch/variants/obj_$NAME; mkdir -p $OUT
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -Xcudafe --diag_suppress=declared_but_not_referenced $EXTRA"
for f in svsb200 search_f32 search_f16 search_i8 search_u8 search_lvq8 fast_f32 fast_f16 fast_i8 fast_u8 fast_lvq8; do
  ( /usr/local/cuda/bin/nvcc $FLAGS -c $f.cu -o $OUT/$f.o ) &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../scratch/variants/libsvsb200_$NAME.so $OUT/*.o build.o build_f32.o build_f16.o flat.o
rm -rf $OUT
echo built libsvsb200_$NAME.so
