#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/ori_bins_diag.py > gpurun_out/ori_bins_diag.txt 2>&1; echo "diag rc=$?"; cat gpurun_out/ori_bins_diag.txt | head -40
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
