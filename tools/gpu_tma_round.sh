timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tma or group_sizes or dense_and_indexed or reproducible" > gpurun_out/r1j_tests.log 2>&1; tail -6 gpurun_out/r1j_tests.log
PROF_READ_REF=1 timeout 300 python tools/prof_rhs_large.py 0,1,1:diag=2,1:diag=1,1:kc=256:st=6,1:kc=384:st=4,1:kc=1024:st=1,1:kc=512:st=2 > gpurun_out/r1j_prof.log 2>&1; tail -10 gpurun_out/r1j_prof.log
STB_TMA=1 timeout 200 python tools/run_c4.py > gpurun_out/r1j_c4_tma.json 2> gpurun_out/r1j_c4_tma.err; cat gpurun_out/r1j_c4_tma.json | cut -c1-900
STB_TMA=0 timeout 200 python tools/run_c4.py > gpurun_out/r1j_c4_reg.json 2> gpurun_out/r1j_c4_reg.err; cat gpurun_out/r1j_c4_reg.json | cut -c1-900
