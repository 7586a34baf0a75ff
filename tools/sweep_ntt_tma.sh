#!/bin/bash
# GPU-box helper: NTT pass kernel variants (tile layout x TMA) on the 2^22 x 64 commit
out=gpurun_out/ntt_tma_sweep.txt; : > $out
run() { python bench.py --steps 5 --warmup 3 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   reorder_and_dft', round(d['breakdown_ms']['reorder_and_dft'],3), 'ms  commit', round(d['ms_per_step'],3), 'ms')" >> $out; }
build() { make -C leanmultisig_b200/csrc -j16 EXTRA="$1" build/ntt.o -B > /dev/null 2>&1; touch leanmultisig_b200/csrc/build/*.o; make -C leanmultisig_b200/csrc > /dev/null 2>&1; }
build ""; echo "XOR swizzle (rows 8/16 apart folded in), cp.async loads + st.global stores  [LM_NTT_NO_TMA=1]" >> $out; LM_NTT_NO_TMA=1 run
build "-DNTT_DENSE_TILE"; echo "dense tile, cp.async + st.global [LM_NTT_NO_TMA=1]" >> $out; LM_NTT_NO_TMA=1 run
echo "dense tile, TMA loads + stores (CU_TENSOR_MAP_SWIZZLE_NONE)" >> $out; LM_NTT_TMA_SWIZZLE=none run
timeout 100 env LM_NTT_TMA_SWIZZLE=none python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dft or reorder or commit" 2>&1 | tail -2 >> $out
build "-DNTT_SWIZZLE_TMA128"; echo "TMA-128B-pattern swizzle, cp.async + st.global [LM_NTT_NO_TMA=1]" >> $out; LM_NTT_NO_TMA=1 run
cat $out
