#!/bin/bash
# Source-level (per SASS instruction) ncu capture of one launch of the stencil kernels: executed counts and stall samples.
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --batch 16 --no-cpu-baseline --no-graph --no-extras"
cap() { name=$1; shift
  ncu --set full --import-source on --clock-control none --kernel-name-base demangled -o gpurun_out/$name -f "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page source --csv > gpurun_out/$name.source.csv 2>/dev/null
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.csv 2>/dev/null
  rm -f gpurun_out/$name.ncu-rep
}
cap src_blur -k "regex:blur_kernel" -s 75 -c 1 $B
cap src_detect -k "regex:detect_rows_kernel" -s 3 -c 1 $B
ls -la gpurun_out | head -30
