#!/bin/bash
mkdir -p gpurun_out
for n in 592 1184 48000; do
AFFNET_B200_LIB=$PWD/affnet_b200/lib/libaffnet_b200_xprof.so timeout 300 python scripts/role_prof_x.py $n 2>&1 | tee gpurun_out/role_prof_x_$n.txt | grep -A3 "slot\|==" | grep -v "^--" | cut -c1-190 | grep "==\|slot\|issuer\|loader"
done
