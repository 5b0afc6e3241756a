#!/bin/bash
# compute-sanitizer over the optional-term kernels (occl_kernel, select_kernel with topk, hypotheses_kernel<true>)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=900 -k "optional or head" 2>&1 | tail -8 > gpurun_out/pytest_f3.log; cat gpurun_out/pytest_f3.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q --timeout=800 -k "optional_score_terms_vs_oracle or opt-" > gpurun_out/sanitizer_f3_memcheck.log 2>&1; echo "memcheck exit $?"; tail -4 gpurun_out/sanitizer_f3_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -x -q --timeout=500 -k "opt-topk or opt-all or opt-occl_" > gpurun_out/sanitizer_f3_racecheck.log 2>&1; echo "racecheck exit $?"; tail -4 gpurun_out/sanitizer_f3_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests -m gpu -x -q --timeout=500 -k "opt-topk or opt-all" > gpurun_out/sanitizer_f3_synccheck.log 2>&1; echo "synccheck exit $?"; tail -4 gpurun_out/sanitizer_f3_synccheck.log
