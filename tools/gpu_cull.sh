set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_seeker_gpu.py tests/test_ops_gpu.py -x -q -m gpu 2>&1 | grep -v "^frame #" > gpurun_out/t7.log; tail -4 gpurun_out/t7.log | cut -c1-300
timeout 300 python tools/stage_times.py --frames 128 > gpurun_out/st_cull.json 2>&1
grep -h '"score"\|"cull"\|run(all' gpurun_out/st_cull.json
timeout 300 python tools/stage_times.py --frames 16 --config cfg5 > gpurun_out/st_cull5.json 2>&1
grep -h '"score"\|"cull"\|run(all' gpurun_out/st_cull5.json
