#!/bin/bash
# GPU-box helper: rebuild ntt.cu with different CTA shapes and time reorder_and_dft at the BASELINE shape.
cd leanmultisig_b200/csrc
for CFG in "512 2" "256 3" "256 4" "384 2" "384 3" "1024 1"; do
  set -- $CFG
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --cudart static -DNTT_THREADS=$1 -DNTT_MIN_BLOCKS=$2 -Xptxas -v -c ntt.cu -o build/ntt.o 2>&1 | grep -A1 "ntt_pass" | grep -E "registers" | tr '\n' ' '
  nvcc -gencode arch=compute_100a,code=sm_100a --cudart static -shared -o ../lib/libleanmultisig_b200.so build/capi.o build/merkle.o build/ntt.o build/poly.o build/sumcheck.o build/air.o build/gkr.o
  echo "== threads=$1 minblocks=$2"
  (cd ../.. && python tools/quick_time.py 28 2>&1 | grep -E "reorder")
done
