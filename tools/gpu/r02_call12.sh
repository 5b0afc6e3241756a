#!/bin/bash
# round 2, call 12: branch-free sweep kernel (unconditional REDs at per-lane scratch words, per-lane queue pushes,
# far-point padding instead of live predicates) vs the previous kernel (-DFNP_SWEEP_V1)
mkdir -p gpurun_out
run() { tag=$1; shift; ( timeout 600 "$@" ) > gpurun_out/r02p_bench_$tag.json 2> gpurun_out/r02p_bench_$tag.err; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02p_bench_$tag.json") if l.startswith("{")][-1])
    print("$tag", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["stage_ms"].items() if k!="note"})
except Exception as e:
    print("$tag", "FAILED", e)
PY
}
timeout 900 python -m pytest tests/test_seeker_gpu.py -m gpu -x -q > gpurun_out/r02p_pytest_seeker.log 2>&1; tail -3 gpurun_out/r02p_pytest_seeker.log
run v2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
FNP_LIB_PATH=$PWD/build_ab/libfnp_sweepv1.so run v1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run cfg5_v2 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --config cfg5 --frames 16 --distinct 8
FNP_LIB_PATH=$PWD/build_ab/libfnp_sweepv1.so run cfg5_v1 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --config cfg5 --frames 16 --distinct 8
