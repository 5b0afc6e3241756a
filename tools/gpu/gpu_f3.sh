#!/bin/bash
# Row f3 check: parity tests, smoke, stage times and the default bench line (no ncu)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=900 2>&1 | tail -25 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python tools/stage_times.py --frames 128 > gpurun_out/stage_times_128.json 2> gpurun_out/stage_times.err; cat gpurun_out/stage_times_128.json; tail -3 gpurun_out/stage_times.err
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
