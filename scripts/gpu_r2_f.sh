#!/bin/bash
mkdir -p gpurun_out
AFFNET_B200_LIB=$PWD/affnet_b200/lib/libaffnet_b200_xprof.so timeout 300 python scripts/role_prof_x.py 48000 2>&1 | tee gpurun_out/role_prof_x.txt | tail -60
