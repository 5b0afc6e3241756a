#!/bin/bash
# compute-sanitizer over the GPU parity tests (memcheck) and the smoke pipeline (racecheck, synccheck)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q --timeout=1400 -k "not cfg5 and not full_size" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"; tail -4 gpurun_out/sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck exit $?"; tail -3 gpurun_out/sanitizer_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/sanitizer_synccheck.log 2>&1; echo "synccheck exit $?"; tail -3 gpurun_out/sanitizer_synccheck.log
