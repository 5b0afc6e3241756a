#!/bin/bash
# A/B of sweep_score_kernel launch shapes (alternate builds under ab/, FNP_LIB_PATH) on 128 cfg2 frames
for lib in "" ab/libfnp_t128_c8.so ab/libfnp_t128_c6.so ab/libfnp_t192_c5.so; do
  for sp in 1024 512 256; do
    echo -n "lib=${lib:-default} split=$sp  "
    FNP_LIB_PATH=${lib:+$PWD/$lib} python tools/stage_times.py --frames 128 --split-points $sp 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print(d['stages_ms']['score'])"
  done
done
