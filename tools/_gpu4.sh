set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_seeker_gpu.py -x -q -m gpu 2>&1 | grep -v "^frame #" | tail -40 > gpurun_out/t4.log
tail -5 gpurun_out/t4.log
for sp in 512 1024 2048; do
timeout 300 python tools/stage_times.py --frames 128 --score-mode sweep --split-points $sp > gpurun_out/st_sweep_$sp.json 2>&1
done
grep -h '"score"\|split_points' gpurun_out/st_sweep_*.json
