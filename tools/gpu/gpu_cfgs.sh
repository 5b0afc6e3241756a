#!/bin/bash
mkdir -p gpurun_out
python tools/stage_times.py --config cfg1 --frames 256 --distinct 16 2> gpurun_out/st_cfg1.err | tee gpurun_out/stage_times_cfg1.json | python -c "import json,sys; d=json.load(sys.stdin); print('cfg1', d['frames'], d['F'], d['H'], d['sum_P_f'], d['tests'], d['split_points'], {k: round(v,4) for k,v in d['stages_ms'].items()})"; tail -2 gpurun_out/st_cfg1.err
python tools/stage_times.py --config cfg5 --frames 8 --distinct 2 2> gpurun_out/st_cfg5.err | tee gpurun_out/stage_times_cfg5.json | python -c "import json,sys; d=json.load(sys.stdin); print('cfg5', d['frames'], d['F'], d['H'], d['sum_P_f'], d['tests'], d['split_points'], {k: round(v,4) for k,v in d['stages_ms'].items()})"; tail -2 gpurun_out/st_cfg5.err
python bench.py --config cfg1 --frames 256 --distinct 16 --steps 10 --warmup 3 --cpu-sample-frames 16 > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err; cut -c1-1200 gpurun_out/bench_cfg1.json; tail -2 gpurun_out/bench_cfg1.err
