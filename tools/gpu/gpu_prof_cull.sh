#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k 'regex:cull_|cell_table|scan_|stats_' -s 6 -c 6 -o gpurun_out/prof_cull -f python bench.py --steps 1 --warmup 1 --frames 32 --no-cpu-baseline > gpurun_out/ncu_cull.log 2>&1
tail -2 gpurun_out/ncu_cull.log
