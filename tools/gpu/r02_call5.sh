#!/bin/bash
# round 2, call 5: sector table + new ops -- GPU suite, bench, launch list, ops timing
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 ) > gpurun_out/r02e_pytest.log 2>&1
tail -8 gpurun_out/r02e_pytest.log
( timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; tail -c 300 gpurun_out/r02e_bench.json; tail -3 gpurun_out/r02e_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:cull_|cell_table|hypotheses_|plan_items|recall_|score_|sweep_|seg_nms|select_|stats_|write_items|upload_' -c 120 --csv --log-file gpurun_out/r02e_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02e_bench_under_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r02e_launches.csv | tee gpurun_out/r02e_launches_summary.txt | head -20
timeout 900 python tools/ref_gpu_bench.py --out gpurun_out/r02e_reference_gpu.json > gpurun_out/r02e_refbench.log 2>&1; tail -c 300 gpurun_out/r02e_refbench.log
