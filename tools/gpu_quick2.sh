set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_seeker_gpu.py -x -q -m gpu -k "nuscenes or reference_compatible" 2>&1 | grep -v "^frame #" | tail -30 > gpurun_out/t5.log; tail -30 gpurun_out/t5.log | cut -c1-300
