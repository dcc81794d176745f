#!/bin/bash
# GPU box: parity of the nets, then the warp-role cycle table of tc_first2_kernel (developer build with -DAG_ROLE_PROF)
mkdir -p gpurun_out

AFFNET_B200_LIB=$PWD/affnet_b200/lib/libaffnet_b200_prof.so timeout 300 python scripts/role_prof.py ${1:-32000} 2>&1 | tee gpurun_out/role_prof.txt
