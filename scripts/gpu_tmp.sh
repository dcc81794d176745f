#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_ab.sh base _ew48 _ew88 base _ew48 _ew88
