#!/bin/bash
# ncu evidence for profiles/ (keep gpurun_out under 64 MiB: no source import, one launch of every kernel).
#   gpurun --timeout 1500 -- 'bash scripts/gpu_profile.sh'      then here:  python scripts/ncu_summary.py r01
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --batch 16 --no-cpu-baseline --no-graph"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_bench.log 2>&1
# one step's worth of tensor-core launches: 3 x tc_first2 + 10 x tc_conv + 2 x tc_conv_pair + tc_head + 2 x tc_headx = 18 (skip the 3 warm-up steps)
ncu --set full --clock-control none --kernel-name-base demangled -k "regex:tc_first2_kernel|tc_conv_kernel|tc_conv_pair_kernel|tc_head_kernel|tc_headx_kernel" -s 54 -c 18 \
    -o gpurun_out/prof_tc -f $B > gpurun_out/ncu_tc.log 2>&1
# detector + selection + filters of one step
ncu --set full --clock-control none --kernel-name-base demangled -k "regex:detect_warp_kernel|resolve_kernel|select_kernel|shape_filter_kernel" -s 12 -c 4 \
    -o gpurun_out/prof_misc -f $B > gpurun_out/ncu_misc.log 2>&1
# octave-0 blurs and the first of octave 1
ncu --set full --clock-control none --kernel-name-base demangled -k "regex:blur_kernel" -s 75 -c 6 \
    -o gpurun_out/prof_blur -f $B > gpurun_out/ncu_blur.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
ls -la gpurun_out; du -sh gpurun_out
