set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_seeker_gpu.py -x -q -m gpu -k "sweep or full_size or stress" 2>&1 | grep -v "^frame #" | tail -5 | cut -c1-300
for v in a b; do
  if [ $v = b ]; then export FNP_LIB_PATH=$PWD/findnpropagate_b200/libfnp_sm100_b.so; fi
  timeout 300 python tools/stage_times.py --frames 128 > gpurun_out/st_$v.json 2>&1
  grep -h '"score"\|"cull"\|run(all' gpurun_out/st_$v.json
done
unset FNP_LIB_PATH
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_q.json").read().strip().splitlines()[-1])
print("value %.0f ms/step %.2f e2e %.0f (%.2f ms)"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["e2e"]["ms_per_step"]), d["host_ms_per_step"])
PY
