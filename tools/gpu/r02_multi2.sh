#!/bin/bash
# 2 GPUs: weak scaling bench + strong-scaling mode (short) -- exercises the multi-rank code paths
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02f_topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02f_bench_n2.json 2> gpurun_out/r02f_bench_n2.err; tail -c 1200 gpurun_out/r02f_bench_n2.json; tail -5 gpurun_out/r02f_bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --warmup 3 --total-frames 1203 --distinct 64 > gpurun_out/r02f_bench_n2_strong.json 2> gpurun_out/r02f_bench_n2_strong.err; tail -c 800 gpurun_out/r02f_bench_n2_strong.json; tail -5 gpurun_out/r02f_bench_n2_strong.err
