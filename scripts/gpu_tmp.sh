#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_ab.sh base _hint base _hint
B="python bench.py --steps 1 --warmup 3 --batch 16 --no-cpu-baseline --no-graph --no-extras"
cap() { name=$1; shift
  ncu --set full --import-source on --clock-control none --kernel-name-base demangled -o gpurun_out/$name -f "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page source --csv > gpurun_out/$name.source.csv 2>/dev/null
  rm -f gpurun_out/$name.ncu-rep
}
cap src_first -k "regex:tcx_first_kernel" -s 9 -c 1 $B
cap src_conv -k "regex:tcx_conv_kernel" -s 37 -c 1 $B
ls -la gpurun_out/src_*
