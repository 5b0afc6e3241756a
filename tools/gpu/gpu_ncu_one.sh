#!/bin/bash
# one ncu --set full capture (source + SASS counters) of the kernels matching $KREGEX on a 128-frame cfg2 batch
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-sweep_score} -s ${SKIP:-2} -c 1 -o gpurun_out/${OUT:-prof_one} -f python tools/stage_times.py --frames 128 --iters 1 > gpurun_out/ncu_one.log 2>&1
tail -2 gpurun_out/ncu_one.log
