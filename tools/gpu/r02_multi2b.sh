#!/bin/bash
# 2 GPUs: is the per-step time of the multi-rank path the single-GPU one?
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02z_bench_n2.json 2> gpurun_out/r02z_bench_n2.err


python - <<'PY'
import json
for t in ("n2",):
    try:
        d=json.loads([l for l in open("gpurun_out/r02z_bench_%s.json" % t) if l.startswith("{")][-1])
        print(t, "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d.get("ms_per_step_by_rank"), d["host_ms_per_step"]["resident"])
    except Exception as e:
        print(t, "FAILED", e)
PY
