#!/bin/bash
# round 2, call 4: GPU suite on the paged stage 1 + vectorised stats, launch list, full ncu of the three big kernels
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 ) > gpurun_out/r02d_pytest.log 2>&1
tail -6 gpurun_out/r02d_pytest.log
( timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; tail -c 300 gpurun_out/r02d_bench.json; tail -3 gpurun_out/r02d_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:cull_|cell_table|hypotheses_|plan_items|recall_|score_|sweep_|seg_nms|select_|stats_|write_items|upload_' -c 120 --csv --log-file gpurun_out/r02d_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02d_bench_under_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r02d_launches.csv | tee gpurun_out/r02d_launches_summary.txt | head -30
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:cull_kernel|stats_kernel|sweep_score_kernel|hypotheses_kernel' -s 4 -c 4 -o gpurun_out/r02d_prof -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02d_ncu_full.log 2>&1
python tools/ncu_kernels.py gpurun_out/r02d_prof.ncu-rep | tee gpurun_out/r02d_ncu_kernels.txt
