TAG=${1:-r1t}
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_tests.log 2>&1; tail -4 gpurun_out/${TAG}_tests.log
STB_TMA_INT=1 timeout 200 python tools/run_c4.py > gpurun_out/${TAG}_c4_tma.json 2> gpurun_out/${TAG}_c4_tma.err; cut -c1-200 gpurun_out/${TAG}_c4_tma.json
STB_TMA_INT=0 timeout 200 python tools/run_c4.py > gpurun_out/${TAG}_c4_reg.json 2> gpurun_out/${TAG}_c4_reg.err; cut -c1-200 gpurun_out/${TAG}_c4_reg.json
timeout 100 python tools/run_c3.py > gpurun_out/${TAG}_c3.log 2>&1; tail -2 gpurun_out/${TAG}_c3.log | cut -c1-400
