set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_seeker_gpu.py -x -q -m gpu 2>&1 | grep -v "^frame #" > gpurun_out/t6.log; tail -3 gpurun_out/t6.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_score -s 2 -c 1 -o gpurun_out/sweep_r3 -f python tools/stage_times.py --frames 128 --iters 1 > gpurun_out/ncu_sweep.log 2>&1
tail -2 gpurun_out/ncu_sweep.log
