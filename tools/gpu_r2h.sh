#!/bin/bash
TAG=${1:-r2h}
mkdir -p gpurun_out
STB_CL_V=1 timeout 300 python tools/prof_cl_opts.py 3 > gpurun_out/${TAG}_v1.log 2>&1; grep -v "cluster integrator" gpurun_out/${TAG}_v1.log | cut -c1-1400 | tail -4
STB_CL_V=2 timeout 300 python tools/prof_cl_opts.py 3 0 > gpurun_out/${TAG}_v2.log 2>&1; grep -v "cluster integrator" gpurun_out/${TAG}_v2.log | cut -c1-1400 | tail -8
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${TAG}_tests.log
timeout 300 python tools/prof_step_gap.py > gpurun_out/${TAG}_gap.log 2>&1; cat gpurun_out/${TAG}_gap.log | tail -8
