#!/bin/bash
# A/B of library variants: gpu_ab.sh "" _minb2 _minb4 ...   (suffixes of affnet_b200/lib/libaffnet_b200<suffix>.so)
mkdir -p gpurun_out
for v in "$@"; do
  [ "$v" = "base" ] && v=""
  AFFNET_B200_LIB=$PWD/affnet_b200/lib/libaffnet_b200$v.so timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_ab$v.json 2> gpurun_out/bench_ab$v.err; echo "bench$v rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_ab$v.json"))
print("[$v]", round(d["value"],1), "Mpix/s", round(d["ms_per_step"],3), "ms/step", {k:round(v,3) for k,v in d["roofline"]["stages_ms"].items() if "detect" in k or "blur" in k or "select" in k or "tcx" in k})
print("   convs:", [round(v,3) for k,v in d["roofline"]["launches_ms"] if k.startswith("tcx_conv")])
PY
done
