#!/bin/bash
# Final evidence of the round.  part a: both bench arms, ncu launch list of the bench, ncu --set full of the
# relaxation launch of the cluster integrator (full bench shape); part b: tests + ncu --set full of the TMA
# RHS / Jv kernels.  (Two parts: gpurun brings back at most 64 MiB per call.)
TAG=${1:-r02}
PART=${2:-a}
mkdir -p gpurun_out
if [ "$PART" = "a" ]; then
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err; cut -c1-300 gpurun_out/${TAG}_bench.json
timeout -k 5 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > gpurun_out/${TAG}_ncu1.log 2>&1; echo "launches rc=$?"
timeout 120 python tools/prof_cl_full.py > gpurun_out/${TAG}_cl_plain.log 2>&1 && tail -1 gpurun_out/${TAG}_cl_plain.log && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_integrate_cl' -s 1 -c 1 -f -o gpurun_out/${TAG}_cl python tools/prof_cl_full.py > gpurun_out/${TAG}_ncu2.log 2>&1; echo "ncu cl rc=$?"
else
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
timeout 300 python tools/prof_rhs_large.py 1 > gpurun_out/${TAG}_rhs_plain.log 2>&1 && tail -1 gpurun_out/${TAG}_rhs_plain.log && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_op_tma' -s 2 -c 2 -f -o gpurun_out/${TAG}_rhs_jv_tma python tools/prof_rhs_large.py 1 > gpurun_out/${TAG}_ncu3.log 2>&1; echo "ncu tma rc=$?"
fi
ls -la gpurun_out/${TAG}_*.ncu-rep
