#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x > gpurun_out/pytest_det.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_det.log
for v in new old; do
  if [ $v = old ]; then export AG_DETECT_WARP_V1=1; else unset AG_DETECT_WARP_V1; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_det_$v.json 2> gpurun_out/bench_det_$v.err; echo "bench rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_det_$v.json"))
print("$v", round(d["value"],1), "Mpix/s", round(d["ms_per_step"],3), "ms/step", {k:round(v,3) for k,v in d["roofline"]["stages_ms"].items() if "detect" in k or "blur" in k or "select" in k})
PY
done
