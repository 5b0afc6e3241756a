#!/bin/bash
# round 2, call 3: paged stage 1 -- GPU suite, sanitizer on the small cases, bench
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 ) > gpurun_out/r02c_pytest.log 2>&1
tail -15 gpurun_out/r02c_pytest.log
timeout 600 compute-sanitizer --tool memcheck python __graft_entry__.py smoke > gpurun_out/r02c_memcheck.log 2>&1; tail -4 gpurun_out/r02c_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python __graft_entry__.py smoke > gpurun_out/r02c_racecheck.log 2>&1; tail -4 gpurun_out/r02c_racecheck.log
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; tail -c 600 gpurun_out/r02c_bench.json; tail -5 gpurun_out/r02c_bench.err
