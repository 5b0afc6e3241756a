#!/bin/bash
# round 2, call 18: stages 1b + 2a fused into one kernel in fnp_seeker_run (A/B by run-time switch), interior-quantile tests
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 ) > gpurun_out/r02v_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02v_pytest_gpu.log
run() { tag=$1; shift; ( timeout 600 "$@" ) > gpurun_out/r02v_bench_$tag.json 2> gpurun_out/r02v_bench_$tag.err; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02v_bench_$tag.json") if l.startswith("{")][-1])
    print("$tag", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), {k: round(v,3) for k,v in d["stage_ms"].items() if k!="note"})
except Exception as e:
    print("$tag", "FAILED", e)
PY
}
run fused python bench.py --steps 20 --warmup 3 --no-cpu-baseline
run split python bench.py --steps 20 --warmup 3 --no-cpu-baseline --opt fuse_stats_hyp=0
run fused2 python bench.py --steps 20 --warmup 3 --no-cpu-baseline
run split2 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --opt fuse_stats_hyp=0
run cfg1_fused python bench.py --steps 20 --warmup 3 --no-cpu-baseline --config cfg1
run cfg1_split python bench.py --steps 20 --warmup 3 --no-cpu-baseline --config cfg1 --opt fuse_stats_hyp=0
run cfg5_fused python bench.py --steps 6 --warmup 3 --no-cpu-baseline --config cfg5 --frames 16 --distinct 8
run cfg5_split python bench.py --steps 6 --warmup 3 --no-cpu-baseline --config cfg5 --frames 16 --distinct 8 --opt fuse_stats_hyp=0
