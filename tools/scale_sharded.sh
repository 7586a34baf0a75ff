#!/bin/bash
# GPU-box helper: sharded (strong-scaling) bench at the given rank counts with per-phase timings of rank 0
mkdir -p gpurun_out
for n in "$@"; do
  if [ "$n" = "1" ]; then
    python bench.py --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/scale_n$n.json
  else
    LM_SHARD_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2970$n bench.py --gpus $n --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/scale_n$n.json
  fi
  python -c "import json,sys; d=json.load(open('gpurun_out/scale_n$n.json')); print('N=$n', 'ms', round(d['ms_per_step'],3), 'Gelem/s', round(d['value'],2), 'e2e ms', round(d['e2e']['ms_per_step'],3), d.get('phases_ms_rank0_last_step'))"
done
