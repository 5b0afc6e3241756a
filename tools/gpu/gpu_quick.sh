#!/bin/bash
# quick GPU check: parity tests, stage times, e2e timeline probe
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --timeout=900 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python tools/stage_times.py --frames 32 2> gpurun_out/stage_times.err | tee gpurun_out/stage_times_32.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['split_points'], {k: round(v,4) for k,v in d['stages_ms'].items()})"; tail -3 gpurun_out/stage_times.err
python tools/e2e_probe.py > gpurun_out/e2e_probe.txt 2>&1; tail -8 gpurun_out/e2e_probe.txt
