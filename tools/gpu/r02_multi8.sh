#!/bin/bash
# 8 GPUs: weak scaling (rank-distinct frames), BASELINE configs[3] (6019 frames, sharded), configs[4] (cfg5 stress)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02i_topo8.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02i_bench_n8.json 2> gpurun_out/r02i_bench_n8.err; tail -c 600 gpurun_out/r02i_bench_n8.json; tail -3 gpurun_out/r02i_bench_n8.err
timeout 600 $TR --master-port 29522 bench.py --gpus 8 --warmup 3 --total-frames 6019 --distinct 128 > gpurun_out/r02i_bench_n8_cfg4_6019frames.json 2> gpurun_out/r02i_bench_n8_cfg4.err; tail -c 400 gpurun_out/r02i_bench_n8_cfg4_6019frames.json; tail -3 gpurun_out/r02i_bench_n8_cfg4.err
timeout 600 $TR --master-port 29523 bench.py --gpus 8 --steps 6 --warmup 3 --config cfg5 --frames 16 --distinct 8 > gpurun_out/r02i_bench_n8_cfg5.json 2> gpurun_out/r02i_bench_n8_cfg5.err; tail -c 400 gpurun_out/r02i_bench_n8_cfg5.json; tail -3 gpurun_out/r02i_bench_n8_cfg5.err
