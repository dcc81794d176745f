#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_variants.py -m gpu -q -x > gpurun_out/pytest_var.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_var.log
bash scripts/gpu_ab.sh base _wr48 _wr64 base _wr48 _wr64 2>&1 | grep -v convs
