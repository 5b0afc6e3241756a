#!/bin/bash
# quick GPU iteration: parity tests, stage times, bench (no ncu)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --timeout=900 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python tools/stage_times.py --frames 32 > gpurun_out/stage_times_32.json 2> gpurun_out/stage_times.err; cat gpurun_out/stage_times_32.json; tail -3 gpurun_out/stage_times.err
python bench.py --steps ${STEPS:-10} --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
