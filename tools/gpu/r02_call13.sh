#!/bin/bash
# round 2, call 13: profiles of the branch-free sweep kernel: launch list, ncu --set full of the four big kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:cull_|cell_table|hypotheses_|plan_items|recall_|score_|sweep_|seg_nms|select_|stats_|write_items|upload_' -c 120 --csv --log-file gpurun_out/r02q_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02q_bench_under_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r02q_launches.csv | tee gpurun_out/r02q_launches_summary.txt | head -20
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:cull_kernel|stats_kernel|sweep_score_kernel|hypotheses_kernel' -s 4 -c 4 -o gpurun_out/r02q_prof -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02q_ncu_full.log 2>&1
python tools/ncu_kernels.py gpurun_out/r02q_prof.ncu-rep | tee gpurun_out/r02q_ncu_kernels.txt
