#!/bin/bash
# final code: weak scaling at N GPUs (N = first argument) with the per-rank times up to the exchange
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 2956$N bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02zz_bench_n$N.json 2> gpurun_out/r02zz_bench_n$N.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02zz_bench_n$N.json") if l.startswith("{")][-1])
print("n$N value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d.get("ms_per_step_by_rank"), str(d.get("gather_check"))[:40])
PY
