set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_p.json 2> gpurun_out/bench_p.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --pack-threads 12 > gpurun_out/bench_p12.json 2> gpurun_out/bench_p12.err
python - <<'PY'
import json
for n in ("p","p12"):
    d=json.loads(open("gpurun_out/bench_%s.json"%n).read().strip().splitlines()[-1])
    print(n,"value %.0f ms/step %.2f e2e %.0f (%.2f ms)"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["e2e"]["ms_per_step"]), d["host_ms_per_step"]["e2e"])
PY
