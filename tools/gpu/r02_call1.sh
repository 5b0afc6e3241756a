#!/bin/bash
# round 2, call 1: full GPU test suite (incl. the reference head on the GPU), reference timing tables, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt
( time timeout 1500 python -m pytest tests -m gpu -q -s -x 2>&1 ) > gpurun_out/r02a_pytest.log 2>&1
tail -5 gpurun_out/r02a_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02a_smoke.log 2>&1; tail -2 gpurun_out/r02a_smoke.log
timeout 900 python tools/ref_gpu_bench.py --out gpurun_out/r02_reference_gpu.json > gpurun_out/r02a_refbench.log 2>&1; tail -c 600 gpurun_out/r02a_refbench.log
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; tail -c 1500 gpurun_out/r02a_bench.json; tail -5 gpurun_out/r02a_bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --layout rows --no-cpu-baseline > gpurun_out/r02a_bench_rows.json 2> gpurun_out/r02a_bench_rows.err
