#!/bin/bash
# round 2, call 19: final state -- full GPU suite, smoke, bench lines (cfg2 default with CPU baseline, cfg1, cfg5, reference arm),
# launch list + ncu --set full of the four big kernels, sanitizer on smoke
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 ) > gpurun_out/r02z_pytest_gpu.log 2>&1; tail -6 gpurun_out/r02z_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z_smoke.log 2>&1; tail -2 gpurun_out/r02z_smoke.log
timeout 900 python bench.py > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err
timeout 600 python bench.py --config cfg1 --no-cpu-baseline > gpurun_out/r02z_bench_cfg1.json 2> gpurun_out/r02z_bench_cfg1.err
timeout 600 python bench.py --config cfg5 --frames 16 --distinct 8 --steps 10 --no-cpu-baseline > gpurun_out/r02z_bench_cfg5.json 2> gpurun_out/r02z_bench_cfg5.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02z_bench_reference_arm.json 2> gpurun_out/r02z_bench_reference_arm.err; tail -c 400 gpurun_out/r02z_bench_reference_arm.json
python - <<'PY'
import json
for t in ("", "_cfg1", "_cfg5"):
    try:
        d=json.loads([l for l in open("gpurun_out/r02z_bench%s.json" % t) if l.startswith("{")][-1])
        print(t or "cfg2", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), {k: round(v,3) for k,v in d["stage_ms"].items() if k!="note"}, "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"],4), "launches", d["gpu_launches"])
    except Exception as e:
        print(t, "FAILED", e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:cull_|cell_table|hypotheses_|plan_items|recall_|score_|sweep_|seg_nms|select_|stats_|write_items|upload_' -c 120 --csv --log-file gpurun_out/r02z_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02z_bench_under_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r02z_launches.csv | tee gpurun_out/r02z_launches_summary.txt | head -16
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:cull_kernel|stats_kernel|sweep_score_kernel|hypotheses_kernel' -s 4 -c 4 -o gpurun_out/r02z_prof -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02z_ncu_full.log 2>&1
python tools/ncu_kernels.py gpurun_out/r02z_prof.ncu-rep | tee gpurun_out/r02z_ncu_kernels.txt
timeout 600 compute-sanitizer --tool memcheck python __graft_entry__.py smoke > gpurun_out/r02z_memcheck.log 2>&1; tail -2 gpurun_out/r02z_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python __graft_entry__.py smoke > gpurun_out/r02z_racecheck.log 2>&1; tail -2 gpurun_out/r02z_racecheck.log
