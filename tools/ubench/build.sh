#!/bin/bash
# Builds the latency micro-benchmarks for sm_100a; run them on the GPU box with: gpurun -- 'tools/ubench/lat; tools/ubench/chains'
set -e
cd "$(dirname "$0")"
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o lat lat.cu
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o chains chains.cu
