#!/bin/bash
# ncu --set full of the three tcx_first launches (+ one conv) of one small step
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --batch 4 --no-cpu-baseline --no-graph --no-extras"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:tcx_first_kernel" -s 9 -c 3 \
    -o gpurun_out/prof_first -f $B > gpurun_out/ncu_first.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_first.log; ls -la gpurun_out/*.ncu-rep
