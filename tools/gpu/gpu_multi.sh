#!/bin/bash
mkdir -p gpurun_out
N=${N:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "exit code $?"; cat gpurun_out/bench_n$N.json; tail -20 gpurun_out/bench_n$N.err
if [ "${REF:-0}" = "1" ]; then
timeout 400 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
fi
