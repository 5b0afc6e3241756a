#!/bin/bash
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck python __graft_entry__.py smoke > gpurun_out/sanitizer.log 2>&1; tail -30 gpurun_out/sanitizer.log
