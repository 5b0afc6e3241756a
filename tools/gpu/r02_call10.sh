#!/bin/bash
# round 2, call 10: stage-1 tile size A/B (passes per CTA sharing one reservation round)
mkdir -p gpurun_out
run() { tag=$1; shift; ( timeout 600 "$@" ) > gpurun_out/r02l_bench_$tag.json 2> gpurun_out/r02l_bench_$tag.err; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02l_bench_$tag.json") if l.startswith("{")][-1])
    print("$tag", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["stage_ms"].items() if k!="note"})
except Exception as e:
    print("$tag", "FAILED", e)
PY
}
run base python bench.py --steps 10 --warmup 3 --no-cpu-baseline
for v in t2048_l1280_c5 t2048_l2048_c4 t2048_l1792_c4 t4096_l2560_c3; do
  FNP_LIB_PATH=$PWD/build_ab/libfnp_$v.so run $v python bench.py --steps 10 --warmup 3 --no-cpu-baseline
  FNP_LIB_PATH=$PWD/build_ab/libfnp_$v.so run cfg5_$v python bench.py --steps 6 --warmup 3 --no-cpu-baseline --config cfg5 --frames 16 --distinct 8
done
run cfg5_base python bench.py --steps 6 --warmup 3 --no-cpu-baseline --config cfg5 --frames 16 --distinct 8
FNP_LIB_PATH=$PWD/build_ab/libfnp_t2048_l1280_c5.so timeout 900 python -m pytest tests/test_seeker_gpu.py -m gpu -q -x > gpurun_out/r02l_pytest_t2048.log 2>&1; tail -3 gpurun_out/r02l_pytest_t2048.log
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 ) > gpurun_out/r02l_pytest.log 2>&1; tail -4 gpurun_out/r02l_pytest.log
