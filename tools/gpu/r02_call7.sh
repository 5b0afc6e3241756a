#!/bin/bash
# round 2, call 7: GPU suite, sweep occupancy A/B, cfg1 / cfg5 lines, bs=1 head timing
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 ) > gpurun_out/r02h_pytest.log 2>&1; tail -5 gpurun_out/r02h_pytest.log
run() { tag=$1; shift; ( timeout 600 "$@" ) > gpurun_out/r02h_bench_$tag.json 2> gpurun_out/r02h_bench_$tag.err; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02h_bench_$tag.json") if l.startswith("{")][-1])
    print("$tag", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), {k: round(v,3) for k,v in d["stage_ms"].items() if k!="note"}, d["host_ms_per_step"]["resident"])
except Exception as e:
    print("$tag", "FAILED", e)
PY
}
run base python bench.py --steps 10 --warmup 3 --no-cpu-baseline
FNP_LIB_PATH=$PWD/build_ab/libfnp_sweep5.so run sweep5 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run cfg1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config cfg1 --frames 256
run cfg1_1024 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config cfg1 --frames 1024
run cfg5 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --config cfg5 --frames 16 --distinct 8
timeout 900 python tools/ref_gpu_bench.py --out gpurun_out/r02h_reference_gpu.json > gpurun_out/r02h_refbench.log 2>&1; python - <<PY
import json
d=json.load(open("gpurun_out/r02h_reference_gpu.json"))
for h in d["heads"]: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in h.items()})
PY
