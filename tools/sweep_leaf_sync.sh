#!/bin/bash
# GPU-box helper: CTA shape / occupancy / CTA-wide barriers of the leaf sponge, timed at the BASELINE shape.
set -e
cd leanmultisig_b200/csrc
OTHERS=$(ls build/*.o | grep -v merkle.o | tr '\n' ' ')
for CFG in "0 128 3" "0 128 2" "1 128 2" "0 256 1" "1 256 1" "1 256 2" "0 128 4" "0 128 1"; do
  set -- $CFG
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --cudart static -DLEAF_SYNC=$1 -DLEAF_THREADS=$2 -DLEAF_MIN_BLOCKS=$3 -Xptxas -v -c merkle.cu -o build/merkle.o 2>&1 | grep -A2 leaf_sponge | grep -E "registers|spill" | tr '\n' ' '
  nvcc -gencode arch=compute_100a,code=sm_100a --cudart static -shared -o ../lib/libleanmultisig_b200.so $OTHERS build/merkle.o
  echo "== SYNC=$1 THREADS=$2 MIN_BLOCKS=$3"
  (cd ../.. && python tools/quick_time.py 28 2>&1 | grep -E "merkle")
done
