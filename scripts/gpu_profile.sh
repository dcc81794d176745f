#!/bin/bash
# ncu evidence for profiles/: launch list of one bench run + full captures of the top kernels.  gpurun --timeout 1500 -- 'bash scripts/gpu_profile.sh'
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --batch 16 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench.log 2>&1
# the three fused first-two-layer kernels + the 12 remaining conv layers + head of one step (skip the 3 warm-up steps: 16 tc launches each)
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:tc_first2_kernel|tc_conv_kernel|tc_head_kernel" -s 48 -c 16 \
    -o gpurun_out/prof_tc -f python bench.py --steps 1 --warmup 3 --batch 16 --no-cpu-baseline --no-graph > gpurun_out/ncu_tc.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:detect_warp_kernel|blur_kernel|select_kernel|orinet_head_kernel" -s 84 -c 28 \
    -o gpurun_out/prof_stencil -f python bench.py --steps 1 --warmup 3 --batch 16 --no-cpu-baseline --no-graph > gpurun_out/ncu_stencil.log 2>&1
ls -la gpurun_out
