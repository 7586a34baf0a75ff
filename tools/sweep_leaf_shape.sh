#!/bin/bash
# GPU-box helper: rebuild merkle.cu with different CTA shapes of the permutation kernels and run the bench ("T:B" pairs)
mkdir -p gpurun_out
out=gpurun_out/leaf_shape_sweep.txt
make -C leanmultisig_b200/csrc -j16 > /dev/null 2>&1
cd leanmultisig_b200/csrc
for v in "$@"; do
  T="${v%%:*}"; B="${v#*:}"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -O2 --cudart static -DLEAF_THREADS=$T -DLEAF_MIN_BLOCKS=$B -Xptxas -v -c merkle.cu -o build/merkle.o 2>&1 | grep -A1 "leaf_sponge" | grep -E "registers|spill" | tr '\n' ' ' >> ../../$out
  nvcc -gencode arch=compute_100a,code=sm_100a --cudart static -shared -o ../lib/libleanmultisig_b200.so build/*.o
  echo "== LEAF_THREADS=$T LEAF_MIN_BLOCKS=$B" >> ../../$out
  (cd ../.. && python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(' ms_per_step', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['breakdown_ms'].items()}, 'e2e_ms', round(d['e2e']['ms_per_step'],3))" >> $out)
done
cat ../../$out
