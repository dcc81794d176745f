#!/bin/bash
mkdir -p gpurun_out
AFFNET_B200_LIB=$PWD/affnet_b200/lib/libaffnet_b200_mc.so timeout 300 python -m pytest tests/test_gpu_tcx.py -m gpu -q -x > gpurun_out/pytest_tcx.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -5 gpurun_out/pytest_tcx.log
if [ $rc -ne 0 ]; then exit 0; fi
bash scripts/gpu_ab.sh base _mc base _mc
