#!/bin/bash
# round 2, call 16: full GPU suite, smoke, default bench (with the CPU baseline), cfg1 / cfg5 lines, reference arm
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 ) > gpurun_out/r02t_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02t_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02t_smoke.log 2>&1; tail -2 gpurun_out/r02t_smoke.log
timeout 900 python bench.py > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err; tail -c 600 gpurun_out/r02t_bench.json
timeout 600 python bench.py --config cfg1 --no-cpu-baseline > gpurun_out/r02t_bench_cfg1.json 2> gpurun_out/r02t_bench_cfg1.err
timeout 600 python bench.py --config cfg5 --frames 16 --distinct 8 --steps 10 --no-cpu-baseline > gpurun_out/r02t_bench_cfg5.json 2> gpurun_out/r02t_bench_cfg5.err
python - <<'PY'
import json
for t in ("", "_cfg1", "_cfg5"):
    try:
        d=json.loads([l for l in open("gpurun_out/r02t_bench%s.json" % t) if l.startswith("{")][-1])
        print(t or "cfg2", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), {k: round(v,3) for k,v in d["stage_ms"].items() if k!="note"}, "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"],4))
    except Exception as e:
        print(t, "FAILED", e)
PY
