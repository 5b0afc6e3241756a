#!/bin/bash
# round 2, call 21: the rewritten sweep kernel at other occupancies: 5 CTAs/SM (48 registers, no spills; needs splits of 1024
# points to fit shared memory), 3 CTAs/SM (78 registers)
mkdir -p gpurun_out
run() { tag=$1; shift; ( timeout 600 "$@" ) > gpurun_out/r02aa_bench_$tag.json 2> gpurun_out/r02aa_bench_$tag.err; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02aa_bench_$tag.json") if l.startswith("{")][-1])
    print("$tag", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["stage_ms"].items() if k!="note"})
except Exception as e:
    print("$tag", "FAILED", e)
PY
}
run base python bench.py --steps 20 --warmup 3 --no-cpu-baseline
FNP_LIB_PATH=$PWD/build_ab/libfnp_sweep5.so run c5_sp1024 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --split-points 1024
FNP_LIB_PATH=$PWD/build_ab/libfnp_sweep5q320.so run c5q320_sp1024 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --split-points 1024
FNP_LIB_PATH=$PWD/build_ab/libfnp_sweep5q320.so run c5q320_sp1280 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --split-points 1280
FNP_LIB_PATH=$PWD/build_ab/libfnp_sweep5.so run c5_sp2048 python bench.py --steps 20 --warmup 3 --no-cpu-baseline
FNP_LIB_PATH=$PWD/build_ab/libfnp_sweep3.so run c3_sp2048 python bench.py --steps 20 --warmup 3 --no-cpu-baseline
run base_sp1024 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --split-points 1024
FNP_LIB_PATH=$PWD/build_ab/libfnp_sweep5.so run cfg5_c5_sp1024 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --config cfg5 --frames 16 --distinct 8 --split-points 1024
run cfg5_base python bench.py --steps 6 --warmup 3 --no-cpu-baseline --config cfg5 --frames 16 --distinct 8
