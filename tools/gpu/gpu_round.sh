#!/bin/bash
# One full GPU session: parity tests, smoke, bench, ncu launch list + full capture of every kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=900 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python tools/stage_times.py --frames 128 > gpurun_out/stage_times_128.json 2> gpurun_out/stage_times.err; tail -3 gpurun_out/stage_times.err
if [ "${NCU:-1}" = "1" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:cull_|cell_table|hypotheses_|plan_items|recall_|scan_|score_|sweep_|seg_nms|select_|stats_|write_items|upload_' -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/launches.csv
ncu --set full --clock-control none --import-source on -k 'regex:cull_|cell_table|hypotheses_|recall_|scan_|score_|sweep_|seg_nms|select_|stats_' -s 16 -c 13 -o gpurun_out/prof_all -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
python tools/ncu_kernels.py gpurun_out/prof_all.ncu-rep
fi
python bench.py --steps 10 --warmup 3 --config cfg1 --no-cpu-baseline > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err; tail -c 400 gpurun_out/bench_cfg1.json
python bench.py --steps 6 --warmup 3 --config cfg5 --frames 16 --distinct 4 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; tail -c 400 gpurun_out/bench_cfg5.json
python bench.py --steps 10 --warmup 3 --score-mode direct --no-cpu-baseline > gpurun_out/bench_direct.json 2> gpurun_out/bench_direct.err; tail -c 300 gpurun_out/bench_direct.json
python tools/probes/feed_probe.py > gpurun_out/feed_probe.txt 2>&1; ./tools/probes/pack_probe > gpurun_out/pack_probe.txt 2>&1
