#!/bin/bash
# compute-sanitizer over the GPU parity tests (memcheck) and small pipelines (racecheck, synccheck)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q --timeout=1400 -k "not cfg5 and not full_size and not cfg2" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"; tail -4 gpurun_out/sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck exit $?"; tail -3 gpurun_out/sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sweep_smoke.py > gpurun_out/sanitizer_racecheck_sweep.log 2>&1; echo "racecheck sweep exit $?"; tail -3 gpurun_out/sanitizer_racecheck_sweep.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sweep_smoke.py > gpurun_out/sanitizer_synccheck_sweep.log 2>&1; echo "synccheck sweep exit $?"; tail -3 gpurun_out/sanitizer_synccheck_sweep.log
