#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-extras > gpurun_out/bench_2gpu_$tag.json 2> gpurun_out/bench_2gpu_$tag.err; echo "$tag rc=$?"; grep -i "symmetric\|error\|Traceback" gpurun_out/bench_2gpu_$tag.err | head -5
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_2gpu_$tag.json"))
    print("$tag", round(d["value"],1), "Mpix/s", round(d["ms_per_step"],3), "ms/step e2e", round(d["e2e"]["value"],1), d.get("exchange"), d["config"]["parallelism"][:90])
except Exception as e: print("$tag failed", e)
PY
}
run ce AG_EXCHANGE=ce
run nccl AG_EXCHANGE=nccl
run nccl_cta2 AG_EXCHANGE=nccl NCCL_MAX_CTAS=2
