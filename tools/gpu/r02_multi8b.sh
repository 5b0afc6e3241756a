#!/bin/bash
# 8 GPUs: weak scaling after the host-side changes (lazy frame views), 3 vs 4 batches in flight
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02m_bench_n8.json 2> gpurun_out/r02m_bench_n8.err; tail -c 300 gpurun_out/r02m_bench_n8.json; tail -2 gpurun_out/r02m_bench_n8.err
timeout 600 $TR --master-port 29532 bench.py --gpus 8 --steps 10 --warmup 3 --slots 4 > gpurun_out/r02m_bench_n8_slots4.json 2> gpurun_out/r02m_bench_n8_slots4.err; tail -c 300 gpurun_out/r02m_bench_n8_slots4.json
timeout 600 $TR --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02m_bench_n8_steps20.json 2> gpurun_out/r02m_bench_n8_steps20.err; tail -c 300 gpurun_out/r02m_bench_n8_steps20.json
