#!/bin/bash
# N-GPU check of the sharded bench (default 8): copy-engine exchange over symmetric memory at full node width.
N=${1:-8}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-extras > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
grep -i "symmetric\|error\|Traceback" gpurun_out/bench_n$N.err | head -5
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n$N.json"))
print(d["n_gpus"], round(d["value"],1), "Mpix/s", round(d["ms_per_step"],3), "ms/step e2e", round(d["e2e"]["value"],1), d.get("exchange"), d["config"])
PY
