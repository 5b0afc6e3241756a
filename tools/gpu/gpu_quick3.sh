#!/bin/bash
# quick A/B: sweep/direct parity tests, stage times (128 frames), short benches (cfg2, cfg5) without the CPU baseline
mkdir -p gpurun_out
python -m pytest tests/test_seeker_gpu.py -m gpu -q --timeout=900 -x -k "${K:-sweep or full_size or cfg5}" 2>&1 | tail -5
python tools/stage_times.py --frames 128 2> gpurun_out/stage_times.err | python -c "import json,sys; d=json.load(sys.stdin); print(json.dumps(d['stages_ms']))"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg2', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_launch'])"
python bench.py --steps 6 --warmup 3 --config cfg5 --frames 16 --distinct 4 --no-cpu-baseline 2> gpurun_out/bench_cfg5.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg5', d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'])"
