#!/bin/bash
mkdir -p gpurun_out
AFFNET_B200_LIB=$PWD/affnet_b200/lib/libaffnet_b200_aew16.so timeout 600 python -m pytest tests/test_gpu_tcx.py -m gpu -q -x > gpurun_out/pytest_tcx.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_tcx.log
bash scripts/gpu_ab.sh base _hew16 _aew16 base _hew16 _aew16
