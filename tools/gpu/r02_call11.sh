#!/bin/bash
# round 2, call 11: stage-1 membership loop, cameras-outer (matrix loads shared by the four sub-tiles) vs sub-tile-outer
mkdir -p gpurun_out
run() { tag=$1; shift; ( timeout 600 "$@" ) > gpurun_out/r02n_bench_$tag.json 2> gpurun_out/r02n_bench_$tag.err; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02n_bench_$tag.json") if l.startswith("{")][-1])
    print("$tag", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["stage_ms"].items() if k!="note"})
except Exception as e:
    print("$tag", "FAILED", e)
PY
}
run subouter python bench.py --steps 10 --warmup 3 --no-cpu-baseline
FNP_LIB_PATH=$PWD/build_ab/libfnp_camouter.so run camouter python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run cfg5_subouter python bench.py --steps 6 --warmup 3 --no-cpu-baseline --config cfg5 --frames 16 --distinct 8
FNP_LIB_PATH=$PWD/build_ab/libfnp_camouter.so run cfg5_camouter python bench.py --steps 6 --warmup 3 --no-cpu-baseline --config cfg5 --frames 16 --distinct 8
run cfg1_subouter python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config cfg1
FNP_LIB_PATH=$PWD/build_ab/libfnp_camouter.so run cfg1_camouter python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config cfg1
FNP_LIB_PATH=$PWD/build_ab/libfnp_camouter.so timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02n_pytest_camouter.log 2>&1; tail -3 gpurun_out/r02n_pytest_camouter.log
