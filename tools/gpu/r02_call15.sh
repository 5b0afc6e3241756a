#!/bin/bash
# round 2, call 15: sweep kernel -- split sizes and warp-queue sizes
mkdir -p gpurun_out
run() { tag=$1; shift; ( timeout 600 "$@" ) > gpurun_out/r02s_bench_$tag.json 2> gpurun_out/r02s_bench_$tag.err; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02s_bench_$tag.json") if l.startswith("{")][-1])
    print("$tag", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["stage_ms"].items() if k!="note"})
except Exception as e:
    print("$tag", "FAILED", e)
PY
}
for sp in 512 1024 1536 2048; do
run sp$sp python bench.py --steps 10 --warmup 3 --no-cpu-baseline --split-points $sp
done
FNP_LIB_PATH=$PWD/build_ab/libfnp_q320.so run q320 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
FNP_LIB_PATH=$PWD/build_ab/libfnp_q512.so run q512 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
FNP_LIB_PATH=$PWD/build_ab/libfnp_q320.so run q320_sp2048 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --split-points 2048
run cfg5_sp2048 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --config cfg5 --frames 16 --distinct 8 --split-points 2048
