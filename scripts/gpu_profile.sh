#!/bin/bash
# ncu evidence for profiles/ (gpurun_out must stay under 64 MiB: every .ncu-rep is exported to its raw-metrics CSV on the box and deleted).
#   gpurun --timeout 2400 -- 'bash scripts/gpu_profile.sh'      then here:  python scripts/ncu_summary.py r02
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/prof_*.csv
cap() {  # cap <name> <ncu args...> -- <command...>
  name=$1; shift
  ncu --set full --clock-control none --kernel-name-base demangled -o gpurun_out/$name -f "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.csv 2>/dev/null
  rm -f gpurun_out/$name.ncu-rep
}
B="python bench.py --steps 1 --warmup 3 --batch 16 --no-cpu-baseline --no-graph --no-extras"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench.log 2>&1
# one step's worth of tensor-core launches: 3 x tcx_first + 12 x tcx_conv + tc_head + 2 x tc_headx = 18 (skip the 3 warm-up steps)
cap prof_tc -k "regex:tcx_first_kernel|tcx_conv_kernel|tc_head_kernel|tc_headx_kernel" -s 54 -c 18 $B
# detector + selection + filters of one step
cap prof_misc -k "regex:detect_rows_kernel|resolve_kernel|select_kernel|shape_filter_kernel" -s 12 -c 4 $B
# octave-0 blurs and the first of octave 1
cap prof_blur -k "regex:blur_kernel" -s 75 -c 6 $B
# BASELINE.json configs[2] (1920x1080, K=4000; 16 of its 64 images per step to bound the capture time): launch list of one step and the
# stencil kernels (octave-0 blurs + detector)
C3="python bench.py --config 3 --batch 16 --steps 1 --warmup 3 --no-cpu-baseline --no-graph --no-extras"
ncu --metrics gpu__time_duration.sum --clock-control none -s 156 -c 52 --csv --log-file gpurun_out/launches_c3.csv $C3 > gpurun_out/ncu_bench_c3.log 2>&1
cap prof_c3_blur -k "regex:blur_kernel" -s 87 -c 5 $C3
cap prof_c3_detect -k "regex:detect_rows_kernel" -s 3 -c 1 $C3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
ls -la gpurun_out; du -sh gpurun_out
