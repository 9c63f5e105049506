#!/bin/bash
TAG=${1:-r2d}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/${TAG}_bench.err; cut -c1-3000 gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err; echo "ref rc=$?"; tail -3 gpurun_out/${TAG}_ref.err; cut -c1-600 gpurun_out/${TAG}_ref.json
