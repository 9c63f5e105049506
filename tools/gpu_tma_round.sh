TAG=${1:-r1l}
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_tests.log 2>&1; tail -6 gpurun_out/${TAG}_tests.log
PROF_READ_REF=1 timeout 300 python tools/prof_rhs_large.py 0,1,1:diag=2,1:diag=1,1:kc=576:st=3,1:kc=384:st=4,1:kc=256:st=6,1:kc=512:st=2 > gpurun_out/${TAG}_prof.log 2>&1; tail -10 gpurun_out/${TAG}_prof.log
STB_LIB=build/lib_v1.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "group_sizes or reproducible" > gpurun_out/${TAG}_tests_v1.log 2>&1; tail -3 gpurun_out/${TAG}_tests_v1.log
