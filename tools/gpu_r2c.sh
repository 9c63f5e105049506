#!/bin/bash
TAG=${1:-r2c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${TAG}_tests.log
timeout 300 python tools/prof_e2e2.py > gpurun_out/${TAG}_e2e.log 2>&1; echo "e2e rc=$?"; tail -8 gpurun_out/${TAG}_e2e.log | cut -c1-700
STB_TIMING=1 timeout 300 python tools/prof_e2e2.py 2>&1 | tail -12 | cut -c1-400 > gpurun_out/${TAG}_e2e_timing.log; cat gpurun_out/${TAG}_e2e_timing.log
