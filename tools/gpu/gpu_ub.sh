#!/bin/bash
mkdir -p gpurun_out
./tools/ubench/score_mix | tee gpurun_out/score_mix.txt
ncu --set full --clock-control none -k regex:kern -s 13 -c 1 -o gpurun_out/prof_ubench -f ./tools/ubench/score_mix > gpurun_out/ncu_ub.log 2>&1; tail -2 gpurun_out/ncu_ub.log
