#!/bin/bash
# round 2, call 8: TMA-staged sweep kernel -- GPU suite, A/B, launch list, full ncu of the four big kernels, sanitizer
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 ) > gpurun_out/r02j_pytest.log 2>&1; tail -5 gpurun_out/r02j_pytest.log
run() { tag=$1; shift; ( timeout 600 "$@" ) > gpurun_out/r02j_bench_$tag.json 2> gpurun_out/r02j_bench_$tag.err; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02j_bench_$tag.json") if l.startswith("{")][-1])
    print("$tag", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), {k: round(v,3) for k,v in d["stage_ms"].items() if k!="note"}, d["host_ms_per_step"]["resident"])
except Exception as e:
    print("$tag", "FAILED", e)
PY
}
run tma1 python bench.py --steps 10 --warmup 3
run tma0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --opt sweep_tma=0
run cfg5_tma1 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --config cfg5 --frames 16 --distinct 8
run cfg5_tma0 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --config cfg5 --frames 16 --distinct 8 --opt sweep_tma=0
timeout 600 compute-sanitizer --tool memcheck python __graft_entry__.py smoke > gpurun_out/r02j_memcheck.log 2>&1; tail -3 gpurun_out/r02j_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python __graft_entry__.py smoke > gpurun_out/r02j_racecheck.log 2>&1; tail -3 gpurun_out/r02j_racecheck.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:cull_|cell_table|hypotheses_|plan_items|recall_|score_|sweep_|seg_nms|select_|stats_|write_items|upload_' -c 120 --csv --log-file gpurun_out/r02j_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02j_bench_under_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r02j_launches.csv | tee gpurun_out/r02j_launches_summary.txt | head -20
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:cull_kernel|stats_kernel|sweep_score_kernel|hypotheses_kernel' -s 4 -c 4 -o gpurun_out/r02j_prof -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02j_ncu_full.log 2>&1
python tools/ncu_kernels.py gpurun_out/r02j_prof.ncu-rep | tee gpurun_out/r02j_ncu_kernels.txt
