#!/bin/bash
# round 2, call 9: warp-autonomous stage 1 -- GPU suite, A/B, sanitizer
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 ) > gpurun_out/r02k_pytest.log 2>&1; tail -5 gpurun_out/r02k_pytest.log
run() { tag=$1; shift; ( timeout 600 "$@" ) > gpurun_out/r02k_bench_$tag.json 2> gpurun_out/r02k_bench_$tag.err; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02k_bench_$tag.json") if l.startswith("{")][-1])
    print("$tag", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), {k: round(v,3) for k,v in d["stage_ms"].items() if k!="note"}, d["host_ms_per_step"]["resident"])
except Exception as e:
    print("$tag", "FAILED", e)
PY
}
run warp1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run warp0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --opt cull_warp=0
run cfg5_warp1 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --config cfg5 --frames 16 --distinct 8
run cfg5_warp0 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --config cfg5 --frames 16 --distinct 8 --opt cull_warp=0
run cfg1_warp1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config cfg1
run cfg1_warp0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config cfg1 --opt cull_warp=0
timeout 600 compute-sanitizer --tool memcheck python __graft_entry__.py smoke > gpurun_out/r02k_memcheck.log 2>&1; tail -3 gpurun_out/r02k_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python __graft_entry__.py smoke > gpurun_out/r02k_racecheck.log 2>&1; tail -3 gpurun_out/r02k_racecheck.log
