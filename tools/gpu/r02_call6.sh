#!/bin/bash
# round 2, call 6: A/B of the sweep variant and of launch shapes (stats / cull), parity of the variants
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 ) > gpurun_out/r02g_pytest.log 2>&1; tail -4 gpurun_out/r02g_pytest.log
run() { tag=$1; shift; ( timeout 600 "$@" ) > gpurun_out/r02g_bench_$tag.json 2> gpurun_out/r02g_bench_$tag.err; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02g_bench_$tag.json") if l.startswith("{")][-1])
    print("$tag", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["stage_ms"].items() if k!="note"})
except Exception as e:
    print("$tag", "FAILED", e)
PY
}
run base python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run sweepv1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --opt sweep_variant=1
for v in stats3 cull4 cull6 list768 cull6l768; do
  FNP_LIB_PATH=$PWD/build_ab/libfnp_$v.so run $v python bench.py --steps 10 --warmup 3 --no-cpu-baseline
done
# parity of sweep variant 1: the sweep-mode tests with the variant switched on
FNP_TEST_SWEEP_VARIANT=1 timeout 900 python -m pytest tests/test_seeker_gpu.py -m gpu -q -x -k "sweep or cfg2 or cfg5 or golden" > gpurun_out/r02g_pytest_sweepv1.log 2>&1; tail -3 gpurun_out/r02g_pytest_sweepv1.log
