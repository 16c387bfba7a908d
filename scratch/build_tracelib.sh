#!/bin/bash
# debug build of the library with the GEMM phase trace (scratch/chain_trace.py); never used by the product
set -e
cd "$(dirname "$0")/.."
P=links-3d-human-pose-estimation_b200
mkdir -p scratch/tracelib
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
/usr/local/cuda/bin/nvcc $F -DLINKS_GEMM_TRACE -c $P/csrc/gemm.cu -o scratch/tracelib/gemm.o
/usr/local/cuda/bin/nvcc -shared -o scratch/tracelib/liblinks_b200.so scratch/tracelib/gemm.o $P/links_b200/_lib/api.o -lcudart
