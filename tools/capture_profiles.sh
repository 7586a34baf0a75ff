#!/bin/bash
# GPU-box helper: launch list of the default bench command + `ncu --set full` captures of the dominant kernels.
#   tools/capture_profiles.sh [tag]     -> gpurun_out/<tag>_launches.csv, <tag>_commit_kernels.ncu-rep (+ summary),
#                                           <tag>_config3_kernels.ncu-rep (+ summary)
tag=${1:-r02}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/${tag}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"leaf_sponge_kernel|ntt_pass_kernel|tree_level_kernel" -c 5 -o gpurun_out/${tag}_commit_kernels -f python bench.py --steps 1 --warmup 1 --no-extras > gpurun_out/${tag}_ncu_full.log 2>&1
python tools/ncu_summary.py gpurun_out/${tag}_commit_kernels.ncu-rep > gpurun_out/${tag}_ncu_commit_kernels.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:"air_exec_round|gkr_round_kernel|gkr_tail_kernel|air_stream_round" -c 14 -o gpurun_out/${tag}_config3_kernels -f python tools/prof_small.py 22 22 > gpurun_out/${tag}_ncu_config3.log 2>&1
python tools/ncu_summary.py gpurun_out/${tag}_config3_kernels.ncu-rep > gpurun_out/${tag}_ncu_air_gkr.txt 2>&1
tail -n 40 gpurun_out/${tag}_ncu_commit_kernels.txt
