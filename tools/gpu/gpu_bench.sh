#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --timeout=900 2>&1 | tail -5
python tools/e2e_probe.py > gpurun_out/e2e_probe.txt 2>&1; tail -4 gpurun_out/e2e_probe.txt
python bench.py --steps ${STEPS:-10} --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
