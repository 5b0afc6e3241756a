set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_seeker_gpu.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/t3.log
cat gpurun_out/t3.log
for sp in 512 1024 2048; do
timeout 300 python tools/stage_times.py --frames 128 --score-mode sweep --split-points $sp > gpurun_out/st_sweep_$sp.json 2>&1
done
grep -h '"score"\|split_points' gpurun_out/st_sweep_*.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_score -s 2 -c 1 -o gpurun_out/sweep_r2 python tools/stage_times.py --frames 128 --score-mode sweep --split-points 1024 --iters 1 > gpurun_out/ncu_sweep.log 2>&1
tail -2 gpurun_out/ncu_sweep.log
