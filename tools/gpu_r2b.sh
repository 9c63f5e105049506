#!/bin/bash
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -30 gpurun_out/${TAG}_tests.log
timeout 300 python tools/prof_e2e2.py > gpurun_out/${TAG}_e2e.log 2>&1; echo "e2e rc=$?"; tail -8 gpurun_out/${TAG}_e2e.log | cut -c1-600
