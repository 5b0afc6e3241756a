#!/bin/bash
# round 2, call 22: default bench lines with every frame of a batch distinct (--distinct 256)
mkdir -p gpurun_out
python bench.py > gpurun_out/r02zz_bench.json 2>/dev/null
python bench.py --config cfg1 --no-cpu-baseline > gpurun_out/r02zz_bench_cfg1.json 2>/dev/null
python - <<'PY'
import json
for t in ("", "_cfg1"):
    d=json.loads([l for l in open("gpurun_out/r02zz_bench%s.json" % t) if l.startswith("{")][-1])
    print(t or "cfg2", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), {k: round(v,3) for k,v in d["stage_ms"].items() if k!="note"}, "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"],4), d["config"]["distinct_frames"])
PY
