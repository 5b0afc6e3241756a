set -x
mkdir -p gpurun_out
for fr in 64 256; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --frames $fr > gpurun_out/bench_f$fr.json 2> gpurun_out/bench_f$fr.err
done
python - <<'PY'
import json
for fr in (64,256):
    d=json.loads(open("gpurun_out/bench_f%d.json"%fr).read().strip().splitlines()[-1])
    print(fr, "value %.0f ms/step %.2f e2e %.0f (%.2f ms)"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["e2e"]["ms_per_step"]), d["host_ms_per_step"]["resident"])
PY
