"""profiles/<tag>_step_launches.md + profiles/<tag>_bench_latest.json from a bench.py JSON line (default gpurun_out/bench_final.json).

    python scripts/bench_to_md.py r01 [path]
"""
import json
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "bench_final.json")
    txt = open(path).read().strip()
    try:
        d = json.loads(txt)                       # an already stored (indented) bench line
    except json.JSONDecodeError:
        d = json.loads(txt.splitlines()[-1])      # raw bench.py output: the JSON line is the last one
    dst = os.path.join(ROOT, "profiles", tag + "_bench_latest.json")
    if os.path.abspath(path) != os.path.abspath(dst):
        json.dump(d, open(dst, "w"), indent=1)
    r = d["roofline"]
    md = ["# %s: one step, kernel by kernel (CUDA events after every launch, `bench.py`)" % tag, "",
          "Workload: %s.  Device-resident %.1f Mpix/s (%.2f ms/step), end to end %.1f Mpix/s; clocks %s." % (
              d["config"]["workload"], d["value"], d["ms_per_step"], d["e2e"]["value"], json.dumps(d["clocks"])), "",
          "| # | kernel | ms |", "|---|---|---|"]
    for i, (k, ms) in enumerate(r["launches_ms"]):
        md.append("| %d | `%s` | %.4f |" % (i + 1, k, ms))
    md += ["", "| kernel family | ms/step |", "|---|---|"]
    for k, ms in r["stages_ms"].items():
        md.append("| `%s` | %.4f |" % (k, ms))
    md += ["", "Tensor-core family: %.2f ms/step = %.1f %% of the step, %.0f TFLOP/s algorithmic = %.1f %% of the %.0f TFLOP/s sustained bf16 peak." % (
        r["kernel_ms_per_step"], 100 * r["share_of_step"], r["achieved"], 100 * r["frac"], r["peak"]),
        "Stencil stage: %.3f ms/step, %.0f GB/s algorithmic = %.1f %% of the %.0f GB/s copy peak." % (
            r["stencil"]["ms_per_step"], r["stencil"]["achieved"], 100 * r["stencil"]["frac"], r["stencil"]["peak"]), ""]
    open(os.path.join(ROOT, "profiles", tag + "_step_launches.md"), "w").write("\n".join(md))
    print("\n".join(md[-4:]))


if __name__ == "__main__":
    main()
