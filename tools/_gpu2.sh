set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_score -s 2 -c 1 -o gpurun_out/sweep_r1 python tools/stage_times.py --frames 128 --score-mode sweep --split-points 1024 --iters 1 > gpurun_out/ncu_sweep.log 2>&1
tail -3 gpurun_out/ncu_sweep.log
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/t2.log
cat gpurun_out/t2.log
