#!/bin/bash
# A/B builds of gemm.cu (never used by the product): scratch/variants/<name>/liblinks_b200.so
set -e
cd "$(dirname "$0")/.."
P=links-3d-human-pose-estimation_b200
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
VARS=("el0:-DLINKS_WGRAD_EVICT_LAST=0" "el1:-DLINKS_WGRAD_EVICT_LAST=1")
for v in "${VARS[@]}"; do
  name=${v%%:*}; flags=${v#*:}
  mkdir -p scratch/variants/$name
  /usr/local/cuda/bin/nvcc $F $flags -c $P/csrc/gemm.cu -o scratch/variants/$name/gemm.o &
done
wait
for v in "${VARS[@]}"; do
  name=${v%%:*}
  /usr/local/cuda/bin/nvcc -shared -o scratch/variants/$name/liblinks_b200.so scratch/variants/$name/gemm.o $P/links_b200/_lib/api.o -lcudart 2>/dev/null
done
