#!/bin/bash
# Build an A/B variant of the library with extra -D flags: tools/abtest.sh NAME -DFLAG ...   -> xpng_b200/build/ab_NAME.so
set -e
cd "$(dirname "$0")/../xpng_b200"
name=$1; shift
mkdir -p build
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC "$@" -c csrc/api.cu -o build/ab_$name.o
nvcc -shared -o build/ab_$name.so build/ab_$name.o build/xpng_file.o build/seven.o build/png7.o build/xpng_pool.o -lz -lpthread -lrt
echo built build/ab_$name.so
