#!/bin/bash
# round 2, call 2: the whole GPU suite without -x, reference timing tables
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 ) > gpurun_out/r02b_pytest.log 2>&1
tail -8 gpurun_out/r02b_pytest.log
timeout 900 python tools/ref_gpu_bench.py --out gpurun_out/r02_reference_gpu.json > gpurun_out/r02b_refbench.log 2>&1; tail -c 1500 gpurun_out/r02b_refbench.log
