#!/bin/bash
TAG=${1:-r2a}
mkdir -p gpurun_out
timeout 300 python tools/prof_cl_opts.py 0 1 2 3 > gpurun_out/${TAG}_opts.log 2>&1; echo "opts rc=$?"; grep -v "^\[stb\] cluster" gpurun_out/${TAG}_opts.log | cut -c1-1500
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
timeout 600 python bench.py --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-1200 gpurun_out/${TAG}_bench.json
