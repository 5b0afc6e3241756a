#!/bin/bash
# 8 GPUs after the sweep-kernel rewrite: weak scaling (20 steps), BASELINE configs[3] (6019 frames sharded), configs[4] (cfg5)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02w_bench_n8.json 2> gpurun_out/r02w_bench_n8.err; tail -c 300 gpurun_out/r02w_bench_n8.json; tail -2 gpurun_out/r02w_bench_n8.err
timeout 600 $TR --master-port 29542 bench.py --gpus 8 --warmup 3 --total-frames 6019 --distinct 128 --no-cpu-baseline > gpurun_out/r02w_bench_n8_cfg4_6019frames.json 2> gpurun_out/r02w_bench_n8_cfg4.err; tail -c 300 gpurun_out/r02w_bench_n8_cfg4_6019frames.json; tail -2 gpurun_out/r02w_bench_n8_cfg4.err
timeout 600 $TR --master-port 29543 bench.py --gpus 8 --steps 6 --warmup 3 --config cfg5 --frames 16 --distinct 8 --no-cpu-baseline > gpurun_out/r02w_bench_n8_cfg5.json 2> gpurun_out/r02w_bench_n8_cfg5.err; tail -c 300 gpurun_out/r02w_bench_n8_cfg5.json; tail -2 gpurun_out/r02w_bench_n8_cfg5.err
python - <<'PY'
import json
for t in ("n8", "n8_cfg4_6019frames", "n8_cfg5"):
    try:
        d=json.loads([l for l in open("gpurun_out/r02w_bench_%s.json" % t) if l.startswith("{")][-1])
        print(t, "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "gather_check", str(d.get("gather_check"))[:60])
    except Exception as e:
        print(t, "FAILED", e)
PY
