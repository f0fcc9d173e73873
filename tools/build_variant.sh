#!/bin/bash
# tools/build_variant.sh NAME SRC.cu [-DFLAG ...]: libfnerf variant with one source recompiled under extra flags
set -e
cd "$(dirname "$0")/.."
NAME=$1; SRC=$2; shift 2
mkdir -p variants
OBJS=$(ls fashion_nerf_b200/build/*.o | grep -v "/${SRC%.cu}.o")
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c fashion_nerf_b200/csrc/$SRC -o /tmp/variant_$NAME.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o variants/libfnerf_$NAME.so $OBJS /tmp/variant_$NAME.o -lcudart
echo variants/libfnerf_$NAME.so
