timeout 2200 compute-sanitizer --tool racecheck python -m pytest tests/test_kitti_gpu.py tests/test_seeker_gpu.py -m gpu -q -x -k "kitti_fused_stages_bit_exact or interior_depth_quantiles and cfg1 or optional_score_terms_vs_oracle" > gpurun_out/r02z_racecheck_suite.log 2>&1; echo rc=$?; tail -3 gpurun_out/r02z_racecheck_suite.log
timeout 600 compute-sanitizer --tool racecheck python tools/sweep_smoke.py cfg2 > gpurun_out/r02z_racecheck_sweep_cfg2.log 2>&1; tail -2 gpurun_out/r02z_racecheck_sweep_cfg2.log
python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r02zz_bench.json 2>/dev/null
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02zz_bench.json") if l.startswith("{")][-1]); print("bench", round(d["value"]), round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["stage_ms"].items() if k!="note"})
PY
