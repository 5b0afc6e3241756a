#!/bin/bash
# GPU iteration: parity tests, stage times, bench, launch list of a 32-frame step (no full ncu capture)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --timeout=900 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python tools/stage_times.py --frames 32 > gpurun_out/stage_times_32.json 2> gpurun_out/stage_times.err; cat gpurun_out/stage_times_32.json; tail -3 gpurun_out/stage_times.err
python bench.py --steps ${STEPS:-10} --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:cull_|hypotheses_|plan_items|recall_|scan_|score_|seg_nms|select_|stats_|write_items' -c 200 --csv --log-file gpurun_out/launches32.csv python bench.py --steps 2 --warmup 1 --frames 32 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/launches32.csv
