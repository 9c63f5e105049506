#!/bin/bash
STB_CLUSTER=1 timeout 300 python tools/prof_midsize.py 1 2>&1 | grep -E "N_uc|default" 
timeout 300 python tools/prof_c3.py 4 2>&1 | grep "streams 16" | cut -c1-200
timeout 300 python tools/prof_cl_opts.py 7 2>&1 | grep "OPTS" | cut -c1-200
