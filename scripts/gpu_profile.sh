#!/bin/bash
# ncu evidence for profiles/ (keep gpurun_out under 64 MiB: no source import, one launch of every kernel).
#   gpurun --timeout 1500 -- 'bash scripts/gpu_profile.sh'
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --batch 16 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1
# one step's worth of tensor-core launches (3 x (first2 + 4 conv) + head = 16; skip the 3 warm-up steps)
ncu --set full --clock-control none --kernel-name-base demangled -k "regex:tc_first2_kernel|tc_conv_kernel|tc_head_kernel" -s 48 -c 16 \
    -o gpurun_out/prof_tc -f python bench.py --steps 1 --warmup 3 --batch 16 --no-cpu-baseline --no-graph > gpurun_out/ncu_tc.log 2>&1
# octave-0 blurs + detector + selection + heads of one step
ncu --set full --clock-control none --kernel-name-base demangled -k "regex:detect_warp_kernel|select_kernel|orinet_head_kernel|affnet_head_kernel|shape_filter_kernel" -s 15 -c 5 \
    -o gpurun_out/prof_misc -f python bench.py --steps 1 --warmup 3 --batch 16 --no-cpu-baseline --no-graph > gpurun_out/ncu_misc.log 2>&1
ncu --set full --clock-control none --kernel-name-base demangled -k "regex:blur_kernel" -s 75 -c 9 \
    -o gpurun_out/prof_blur -f python bench.py --steps 1 --warmup 3 --batch 16 --no-cpu-baseline --no-graph > gpurun_out/ncu_blur.log 2>&1
ls -la gpurun_out; du -sh gpurun_out
