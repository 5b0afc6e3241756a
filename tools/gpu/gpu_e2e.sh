mkdir -p gpurun_out
nproc
for i in 1 2; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e$i.json 2> gpurun_out/bench_e$i.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_e$i.json").read().strip().splitlines()[-1])
print("value %.0f ms/step %.2f e2e %.0f (%.2f ms)"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["e2e"]["ms_per_step"]), d["e2e"]["host_pack"])
PY
done
python tools/probes/feed_probe.py 2>&1 | grep threads
