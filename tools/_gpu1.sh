set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_seeker_gpu.py -x -q -m gpu -k "sweep or auto_picks or golden or batch_of" 2>&1 | tail -30 > gpurun_out/t1.log
cat gpurun_out/t1.log
for sp in 512 1024 2048; do
timeout 300 python tools/stage_times.py --frames 128 --score-mode sweep --split-points $sp > gpurun_out/st_sweep_$sp.json 2>&1
done
timeout 300 python tools/stage_times.py --frames 128 --score-mode direct > gpurun_out/st_direct.json 2>&1
grep -h '"score"\|split_points\|score_mode\|"cull"' gpurun_out/st_*.json
