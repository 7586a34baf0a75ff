#!/bin/bash
# A/B of the leaf kernels' grid (one CTA per 256 rows vs persistent grids) and of the commit pipeline's column groups.
mkdir -p gpurun_out
for g in 0 296 592 1184; do
  echo "LM_LEAF_GRID=$g" >> gpurun_out/leaf_grid_sweep.txt
  LM_LEAF_GRID=$g python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(' ms_per_step', round(d['ms_per_step'],3), d['breakdown_ms'], 'e2e_ms', round(d['e2e']['ms_per_step'],3))" >> gpurun_out/leaf_grid_sweep.txt
done
echo "LM_COMMIT_EVEN_GROUPS=1 (old pipeline shape), LM_LEAF_GRID=0" >> gpurun_out/leaf_grid_sweep.txt
LM_COMMIT_EVEN_GROUPS=1 python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(' ms_per_step', round(d['ms_per_step'],3), 'e2e_ms', round(d['e2e']['ms_per_step'],3))" >> gpurun_out/leaf_grid_sweep.txt
cat gpurun_out/leaf_grid_sweep.txt
