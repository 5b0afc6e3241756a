#!/bin/bash
# last check of a session: parity tests + smoke
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=900 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -1
