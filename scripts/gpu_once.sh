#!/bin/bash
mkdir -p gpurun_out
AFFNET_B200_LIB=$PWD/affnet_b200/lib/libaffnet_b200_both8.so timeout 400 python -m pytest tests/test_gpu_tcx.py -m gpu -q -x -s > gpurun_out/pytest_tcx.log 2>&1; rc=$?; echo "pytest tcx(both8) rc=$rc"; grep -E "engine|passed|failed|Error" gpurun_out/pytest_tcx.log | cut -c1-170 | head -5
bash scripts/gpu_ab.sh base _lo16 _both8 base _lo16 _both8
