#!/bin/bash
TAG=${1:-r2o}
mkdir -p gpurun_out
timeout 300 python tools/prof_cl_opts.py 7 3 > gpurun_out/${TAG}_opts.log 2>&1; grep "cluster integrator" gpurun_out/${TAG}_opts.log | head -1; grep -v "cluster integrator" gpurun_out/${TAG}_opts.log | cut -c1-1500 | grep -v "steps per system"
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${TAG}_tests.log
