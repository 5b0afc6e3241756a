#!/bin/bash
# round 2, call 14: stage 1 with an L2 prefetch of the rows of a later tile (distance in tiles), A/B by run-time switch
mkdir -p gpurun_out
run() { tag=$1; shift; ( timeout 600 "$@" ) > gpurun_out/r02r_bench_$tag.json 2> gpurun_out/r02r_bench_$tag.err; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02r_bench_$tag.json") if l.startswith("{")][-1])
    print("$tag", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["stage_ms"].items() if k!="note"})
except Exception as e:
    print("$tag", "FAILED", e)
PY
}
for pf in 0 370 740 1480 2960; do
run pf$pf python bench.py --steps 10 --warmup 3 --no-cpu-baseline --opt cull_prefetch=$pf
done
run cfg5_pf0 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --config cfg5 --frames 16 --distinct 8
run cfg5_pf740 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --config cfg5 --frames 16 --distinct 8 --opt cull_prefetch=740
