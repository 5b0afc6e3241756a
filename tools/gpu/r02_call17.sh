#!/bin/bash
# round 2, call 17: stats kernel (no centre quantile without the distance term, two quantiles per histogram pass), hypotheses
# kernel occupancy variants
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 ) > gpurun_out/r02u_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02u_pytest_gpu.log
run() { tag=$1; shift; ( timeout 600 "$@" ) > gpurun_out/r02u_bench_$tag.json 2> gpurun_out/r02u_bench_$tag.err; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02u_bench_$tag.json") if l.startswith("{")][-1])
    print("$tag", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["stage_ms"].items() if k!="note"})
except Exception as e:
    print("$tag", "FAILED", e)
PY
}
run base python bench.py --steps 10 --warmup 3 --no-cpu-baseline
for n in 5 7 8; do
FNP_LIB_PATH=$PWD/build_ab/libfnp_hyp$n.so run hyp$n python bench.py --steps 10 --warmup 3 --no-cpu-baseline
done
run cfg1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config cfg1
run cfg5 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --config cfg5 --frames 16 --distinct 8
timeout 600 compute-sanitizer --tool racecheck python __graft_entry__.py smoke > gpurun_out/r02u_racecheck.log 2>&1; tail -3 gpurun_out/r02u_racecheck.log
timeout 600 compute-sanitizer --tool memcheck python __graft_entry__.py smoke > gpurun_out/r02u_memcheck.log 2>&1; tail -3 gpurun_out/r02u_memcheck.log
