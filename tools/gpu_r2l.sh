#!/bin/bash
TAG=${1:-r2l}
mkdir -p gpurun_out
STB_CL_V=2 timeout 300 python tools/prof_cl_opts.py 3 > gpurun_out/${TAG}_v2.log 2>&1; grep "cluster integrator" gpurun_out/${TAG}_v2.log | head -1; grep -v "cluster integrator" gpurun_out/${TAG}_v2.log | cut -c1-1600 | tail -3 | head -2
STB_CL_V=1 timeout 300 python tools/prof_cl_opts.py 3 > gpurun_out/${TAG}_v1.log 2>&1; grep "cluster integrator" gpurun_out/${TAG}_v1.log | head -1; grep -v "cluster integrator" gpurun_out/${TAG}_v1.log | cut -c1-1600 | tail -3 | head -2
