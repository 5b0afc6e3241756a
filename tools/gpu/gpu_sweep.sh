#!/bin/bash
mkdir -p gpurun_out
for sp in 256 512 1024 2048; do
python tools/stage_times.py --frames 32 --split-points $sp 2> gpurun_out/stage_times.err | python -c "import json,sys; d=json.load(sys.stdin); print(d['split_points'], {k: round(v,4) for k,v in d['stages_ms'].items()})"
done | tee gpurun_out/sp_sweep.txt
ncu --set full --clock-control none --import-source on -k 'regex:cull_|hypotheses_|recall_|scan_tiles|seg_nms|select_|stats_' -s 9 -c 9 -o gpurun_out/prof_stages -f python bench.py --steps 1 --warmup 1 --frames 32 --no-cpu-baseline > gpurun_out/ncu_stages.log 2>&1
tail -2 gpurun_out/ncu_stages.log
