#!/bin/bash
# quick check of a stage-1 change: seeker parity tests, stage times (128 frames), short bench
mkdir -p gpurun_out
python -m pytest tests/test_seeker_gpu.py -m gpu -q --timeout=900 -x 2>&1 | tail -4
python tools/stage_times.py --frames 128 2> gpurun_out/stage_times.err | python -c "import json,sys; d=json.load(sys.stdin); print(json.dumps(d['stages_ms']))"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg2', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline_second_kernel']['ms_per_launch'])"
