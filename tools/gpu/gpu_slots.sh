mkdir -p gpurun_out
for sl in 2 3 4; do
timeout 600 python bench.py --steps 12 --warmup 4 --no-cpu-baseline --slots $sl > gpurun_out/bench_s$sl.json 2> gpurun_out/bench_s$sl.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_s$sl.json").read().strip().splitlines()[-1])
    print("slots $sl value %.0f ms/step %.2f e2e %.0f (%.2f ms)"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["e2e"]["ms_per_step"]))
except Exception as e:
    print("slots $sl ERR", e); print(open("gpurun_out/bench_s$sl.err").read()[-1500:])
PY
done
