#!/bin/bash
# A/B builds of gemm.cu (never used by the product): scratch/variants/<name>/liblinks_b200.so
set -e
cd "$(dirname "$0")/.."
P=links-3d-human-pose-estimation_b200
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
for v in "cs0_pf0:-DLINKS_ADAM_CS=0 -DLINKS_ADAM_PREFETCH=0" "cs1_pf0:-DLINKS_ADAM_CS=1 -DLINKS_ADAM_PREFETCH=0" "cs0_pf1:-DLINKS_ADAM_CS=0 -DLINKS_ADAM_PREFETCH=1" "cs1_pf1:-DLINKS_ADAM_CS=1 -DLINKS_ADAM_PREFETCH=1"; do
  name=${v%%:*}; flags=${v#*:}
  mkdir -p scratch/variants/$name
  /usr/local/cuda/bin/nvcc $F $flags -c $P/csrc/gemm.cu -o scratch/variants/$name/gemm.o &
done
wait
for name in cs0_pf0 cs1_pf0 cs0_pf1 cs1_pf1; do
  /usr/local/cuda/bin/nvcc -shared -o scratch/variants/$name/liblinks_b200.so scratch/variants/$name/gemm.o $P/links_b200/_lib/api.o -lcudart 2>/dev/null
done
