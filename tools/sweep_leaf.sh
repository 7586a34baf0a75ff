#!/bin/bash
# GPU-box helper: rebuild merkle.cu with different occupancy targets and time the commit stages.
set -e
cd leanmultisig_b200/csrc
for MB in 2 3 4 5 6; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --cudart static -DLEAF_MIN_BLOCKS=$MB -Xptxas -v -c merkle.cu -o build/merkle.o 2>&1 | grep -A1 leaf_sponge | grep -E "registers|spill" | tr '\n' ' '
  nvcc -gencode arch=compute_100a,code=sm_100a --cudart static -shared -o ../lib/libleanmultisig_b200.so build/capi.o build/merkle.o build/ntt.o build/poly.o
  echo "== LEAF_MIN_BLOCKS=$MB"
  (cd ../.. && python tools/quick_time.py 28 2>&1 | grep -E "merkle|poseidon")
done
