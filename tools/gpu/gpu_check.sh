set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^frame #" > gpurun_out/t_final.log; tail -4 gpurun_out/t_final.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_chk.json 2> gpurun_out/bench_chk.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_chk.json").read().strip().splitlines()[-1])
print("value %.0f ms/step %.2f e2e %.0f (%.2f ms)"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["e2e"]["ms_per_step"]), d["host_ms_per_step"]["resident"])
PY
