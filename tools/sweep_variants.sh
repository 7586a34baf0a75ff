#!/bin/bash
# GPU-box helper: rebuild the library with different -D flags and run the bench for each ("name:flags" arguments).
mkdir -p gpurun_out
out=gpurun_out/variant_sweep.txt
for v in "$@"; do
  name="${v%%:*}"; flags="${v#*:}"
  make -C leanmultisig_b200/csrc -B -j16 EXTRA="$flags" > /dev/null 2>&1 || { echo "$name: build failed" >> $out; continue; }
  echo "== $name ($flags)" >> $out
  python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(' ms_per_step', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['breakdown_ms'].items()}, 'e2e_ms', round(d['e2e']['ms_per_step'],3))" >> $out
done
cat $out
