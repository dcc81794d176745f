"""Turns the ncu reports that scripts/gpu_profile.sh leaves in gpurun_out/ into profiles/<tag>_ncu.md and
profiles/<tag>_ncu_traffic.json (DRAM bytes per launch of every tcgen05 kernel: bench.py's roofline.traffic).

    python scripts/ncu_summary.py r01            # reads gpurun_out/prof_{tc,misc,blur}.ncu-rep + launches.csv
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
OUT = os.path.join(ROOT, "gpurun_out")
COLS = [("time", "gpu__time_duration.sum"), ("dram rd", "dram__bytes_read.sum"), ("dram wr", "dram__bytes_write.sum"),
        ("dram %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("tensor %", "sm__inst_executed_pipe_tensor_op_gmma.avg.pct_of_peak_sustained_active|sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active|sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active|sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("sm %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("smem %", "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("L2 hit %", "lts__t_sector_hit_rate.pct"),
        ("regs", "launch__registers_per_thread"), ("occ %", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("grid", "launch__grid_size")]


def raw_rows(rep):
    """rep: a .ncu-rep (exported here with `ncu -i`) or the raw-metrics CSV scripts/gpu_profile.sh exported on the GPU box."""
    if rep.endswith(".csv"):
        txt = open(rep).read()
    else:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    return hdr, units, body


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def short(name):
    name = re.sub(r"^(void )?(ag::)?(tcx?::)?", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("(int)", "")


def table(rep, md, traffic=None):
    hdr, units, body = raw_rows(rep)
    ix = {h: i for i, h in enumerate(hdr)}
    md.append("| kernel | " + " | ".join(c for c, _ in COLS) + " |")
    md.append("|---" * (len(COLS) + 1) + "|")
    for r in body:
        cells = []
        for c, keys in COLS:
            k = next((k for k in keys.split("|") if k in ix), None)
            if k is None:
                cells.append("-")
                continue
            v, u = r[ix[k]], units[ix[k]]
            try:
                v = "%.4g" % float(v.replace(",", ""))
            except ValueError:
                pass
            cells.append(("%s %s" % (v, u)).strip() if c in ("time", "dram rd", "dram wr") else v)
        name = short(r[ix["Kernel Name"]])
        md.append("| `%s` | " % name + " | ".join(cells) + " |")
        if traffic is not None:
            b = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]) + to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
            traffic.append({"kernel": name, "dram_bytes": b})


def launch_table(lp, md, title):
    rows = [r for r in csv.reader(open(lp)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    body = rows[rows.index(hdr) + 1:]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, total = {}, 0.0
    for r in body:
        try:
            ns = float(r[mv].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(r[mu], 1)
        except ValueError:
            continue
        k = short(r[kn])
        a = agg.setdefault(k, [0.0, 0])
        a[0] += ns; a[1] += 1; total += ns
    md += ["## " + title % (sum(a[1] for a in agg.values()), total / 1e6), "", "| ms | share | launches | kernel |", "|---|---|---|---|"]
    for k, (ns, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        md.append("| %.3f | %.1f%% | %d | `%s` |" % (ns / 1e6, 100 * ns / total, n, k))
    md.append("")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    md = ["# %s: `ncu --set full --clock-control none` captures on B200" % tag, "",
          "Command: `scripts/gpu_profile.sh` (bench.py --batch 16 --no-graph, one launch of every kernel of one step after three warm-up steps; every report is "
          "exported to its raw-metrics CSV on the box); this file: `python scripts/ncu_summary.py %s`." % tag,
          "Times under ncu are cold-cache and serialised: use the ratios (DRAM %, tensor %, issue %), not the absolute times.",
          "Template arguments: `tcx_first_kernel<C1,COUT,SA,SW,OSA,BF>` (sampler + layers 1-2), `tcx_conv_kernel<CIN,COUT,H,STRIDE,NSPLIT,STAGES,OUT,SA,SW,OSA,EW,BF>` "
          "(OUT: 1 16x16 stride-1 consumer, 2 / 3 pair layouts of the 8x8 layers, 4 head operand; SA / SW / OSA: activation / weight / output residual planes), "
          "`tc_headx_kernel<0 AffNet | 1 OriNet>`, `tc_head_kernel<BF>` (HardNet).", ""]
    traffic = []
    for title, rep, tr in (("tensor-core kernels (AffNet, OriNet, HardNet in launch order)", "prof_tc", traffic),
                           ("detector, selection, filters", "prof_misc", None), ("blur kernels (octave 0 and the first of octave 1 of one step)", "prof_blur", None),
                           ("BASELINE.json configs[2] (1920x1080, K=4000, 16 images per step): octave-0 blurs", "prof_c3_blur", None),
                           ("BASELINE.json configs[2]: fused detector", "prof_c3_detect", None)):
        path = next((os.path.join(OUT, rep + e) for e in (".csv", ".ncu-rep") if os.path.isfile(os.path.join(OUT, rep + e))), None)
        if path is None:
            continue
        md += ["## " + title, ""]
        table(path, md, tr)
        md.append("")
    lp = os.path.join(OUT, "launches.csv")
    if os.path.isfile(lp):
        launch_table(lp, md, "launch list (`ncu --metrics gpu__time_duration.sum`, %d launches = warm-up + timed steps; %.1f ms of kernel time)")
    lp = os.path.join(OUT, "launches_c3.csv")
    if os.path.isfile(lp):
        launch_table(lp, md, "configs[2] launch list: one step of 16 images 1920x1080, K=4000 (%d launches, %.1f ms of kernel time)")
    open(os.path.join(ROOT, "profiles", tag + "_ncu.md"), "w").write("\n".join(md))
    if traffic:
        json.dump({"batch": 16, "source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, one step", "launches": traffic,
                   "family_bytes_per_step": sum(t["dram_bytes"] for t in traffic)}, open(os.path.join(ROOT, "profiles", tag + "_ncu_traffic.json"), "w"), indent=1)
    print("\n".join(md))


if __name__ == "__main__":
    main()
