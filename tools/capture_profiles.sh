#!/bin/bash
# GPU-box helper: launch list of the default bench command + one `ncu --set full` capture of the commit kernels.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"leaf_sponge_kernel|ntt_pass_kernel|tree_level_kernel" -c 5 -o gpurun_out/commit_kernels -f python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_full.log 2>&1
python tools/ncu_summary.py gpurun_out/commit_kernels.ncu-rep > gpurun_out/ncu_commit_kernels_summary.txt 2>&1
tail -n 60 gpurun_out/ncu_commit_kernels_summary.txt
