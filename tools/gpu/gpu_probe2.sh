#!/bin/bash
set -x
mkdir -p gpurun_out
./tools/ubench/pipe_rates > gpurun_out/pipe_rates.txt 2>&1; cat gpurun_out/pipe_rates.txt
python tools/stage_times.py --frames 32 > gpurun_out/stage_times_32.json 2> gpurun_out/stage_times.err; cat gpurun_out/stage_times_32.json; tail -3 gpurun_out/stage_times.err
python tools/stage_times.py --frames 8 > gpurun_out/stage_times_8.json 2>> gpurun_out/stage_times.err
python tools/stage_times.py --config cfg1 --frames 64 > gpurun_out/stage_times_cfg1.json 2>> gpurun_out/stage_times.err
# one full ncu capture of each distinct kernel (first steady-state launch), 8-frame batch
ncu --set full --clock-control none --import-source on -k 'regex:cull_|hypotheses_|recall_|scan_tiles|seg_nms|select_|stats_' -s 12 -c 9 -o gpurun_out/prof_stages -f python bench.py --steps 1 --warmup 1 --frames 8 --no-cpu-baseline > gpurun_out/ncu_stages.log 2>&1
tail -2 gpurun_out/ncu_stages.log
