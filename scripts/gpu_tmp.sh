#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "detect or nms or end_to_end or benchmark" > gpurun_out/pytest_det.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_det.log
bash scripts/gpu_ab.sh base _minb3 base _minb3
