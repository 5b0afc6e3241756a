#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --timeout=900 2>&1 | tail -5
python tools/stage_times.py --frames 32 2> gpurun_out/stage_times.err | tee gpurun_out/stage_times_32.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['split_points'], {k: round(v,4) for k,v in d['stages_ms'].items()})"; tail -3 gpurun_out/stage_times.err
ncu --set full --clock-control none --import-source on -k 'regex:score_kernel|cull_stage' -s 2 -c 2 -o gpurun_out/prof_score2 -f python bench.py --steps 1 --warmup 1 --frames 32 --no-cpu-baseline > gpurun_out/ncu_score2.log 2>&1
tail -2 gpurun_out/ncu_score2.log
