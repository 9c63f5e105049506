#!/bin/bash
# ncu --set full of the relaxation launch (short output grid) of the cluster integrator, v2 and v1
TAG=${1:-r2m}
mkdir -p gpurun_out
for V in 2 1; do
  STB_CL_V=$V timeout 120 python tools/prof_cl_short.py > gpurun_out/${TAG}_plain_v$V.log 2>&1 || { echo "plain run failed"; tail -3 gpurun_out/${TAG}_plain_v$V.log; exit 1; }
  STB_CL_V=$V timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_integrate_cl' -s 1 -c 1 -f -o gpurun_out/${TAG}_cl_v$V python tools/prof_cl_short.py > gpurun_out/${TAG}_ncu_v$V.log 2>&1; echo "ncu v$V rc=$?"
done
ls -la gpurun_out/${TAG}_cl_v*.ncu-rep
