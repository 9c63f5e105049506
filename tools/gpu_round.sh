#!/bin/bash
# One GPU call: parity tests, both bench arms, ncu launch list of the bench, full capture of RHS/Jv (TMA kernels).
TAG=${1:-r1m}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/${TAG}_tests.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json
timeout -k 5 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/${TAG}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_ncu1.log 2>&1; echo "launches rc=$?"
timeout 300 python tools/prof_rhs_large.py 1 > gpurun_out/${TAG}_rhs_plain.log 2>&1 && tail -1 gpurun_out/${TAG}_rhs_plain.log && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_op_tma' -s 2 -c 2 -f -o gpurun_out/${TAG}_rhs_jv_tma python tools/prof_rhs_large.py 1 > gpurun_out/${TAG}_ncu2.log 2>&1; echo "ncu full rc=$?"
