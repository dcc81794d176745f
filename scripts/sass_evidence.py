"""profiles/<tag>_sass_evidence.md: which tcgen05 / TMEM / bulk-copy / cluster SASS instructions every kernel of the built library
contains (cuobjdump -sass; mnemonics per /opt/skills/guides/B200_PROFILING.md).  Runs without a GPU.

    python scripts/sass_evidence.py r01
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
PAT = collections.OrderedDict([
    ("UTC*MMA (tcgen05.mma)", r"\bUTC[A-Z]*MMA\b"), ("LDTM (tcgen05.ld)", r"\bLDTM\b"), ("UTCBAR (tcgen05.commit)", r"\bUTCBAR\b"),
    ("UTCATOM* (tcgen05.alloc/dealloc)", r"\bUTCATOM[A-Z.]*"), ("UBLKCP (cp.async.bulk)", r"\bUBLKCP\b"), ("UTMALDG (TMA tensor map)", r"\bUTMALDG\b"), ("SYNCS (mbarrier)", r"\bSYNCS\b"),
    ("UCGABAR (cluster barrier)", r"\bUCGABAR[A-Z_.]*"), ("LDGSTS (cp.async)", r"\bLDGSTS\b"), ("ELECT", r"\bELECT\b"),
])


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    lib = os.path.join(ROOT, "affnet_b200", "lib", "libaffnet_b200.so")
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    kernels, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            kernels[cur]["_n"] = 0
            continue
        if cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            kernels[cur]["_n"] += 1
            for k, p in PAT.items():
                if re.search(p, line):
                    kernels[cur][k] += 1
    md = ["# %s: SASS evidence (cuobjdump -sass of affnet_b200/lib/libaffnet_b200.so, sm_100a)" % tag, "",
          "Instruction counts per kernel; made by `python scripts/sass_evidence.py %s` (no GPU needed)." % tag, "",
          "| kernel | SASS instr | " + " | ".join(PAT) + " |", "|---" * (len(PAT) + 2) + "|"]
    for k, c in kernels.items():
        name = re.sub(r"^(void )?(ag::)?(tcx?::|pf::)?", "", demangle(k))
        name = re.sub(r"\(.*$", "", name)
        if not any(c[p] for p in list(PAT)[:6]) and not name.startswith(("detect_warp", "detect_rows", "blur", "select_kernel", "octave")):
            continue
        md.append("| `%s` | %d | " % (name, c["_n"]) + " | ".join(str(c[p]) for p in PAT) + " |")
    open(os.path.join(ROOT, "profiles", tag + "_sass_evidence.md"), "w").write("\n".join(md) + "\n")
    print("\n".join(md[:14]))


if __name__ == "__main__":
    main()
