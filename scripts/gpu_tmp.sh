#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x > gpurun_out/pytest_pyr.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_pyr.log
for v in tma notma tma notma; do
  if [ $v = notma ]; then export AG_BLUR_NO_TMA=1; else unset AG_BLUR_NO_TMA; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; echo "bench rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$v.json"))
print("$v", round(d["value"],1), "Mpix/s", round(d["ms_per_step"],3), "ms/step", {k:round(v,3) for k,v in d["roofline"]["stages_ms"].items() if "detect" in k or "blur" in k or "select" in k})
PY
done
