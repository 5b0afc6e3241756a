#!/bin/bash
# parity tests only
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=900 ${PYTEST_ARGS} 2>&1 | tail -${TAIL:-40} > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
