"""Instruction / stall map of the E-step by PHASE from an .ncu-rep (needs -lineinfo, --import-source on).
SASS rows are walked in address order; rows that belong to inlined helpers (dist2, exp_neg, shuffles ...) inherit the phase
of the nearest preceding row that maps to a line of tq_estep_chunk.  Usage: python scripts/ncu_phase_map.py X.ncu-rep"""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
SRC = "trackdlo_b200/csrc/tdlo_taskq.cuh"
# phase boundaries are found from marker comments in the source so that the script follows edits
marks = [("prefetch", "software prefetch"), ("sphere", "---- nearest node"), ("scan", "double best = 1e300"),
         ("neighbours", "whole column underflows"), ("window", "---- node window"), ("phaseA", "double w = 0.0;"),
         ("normalise", "if (quirk) {"), ("phaseB", "---- phase B"), ("epilogue", "per-warp tile-loop cycles"),
         ("END", "// Visibility pre-pass over one chunk")]
lines = open(SRC).read().splitlines()
start = next(i for i, l in enumerate(lines) if "static __device__ void tq_estep_chunk" in l) + 1
bounds = []
for name, pat in marks:
    ln = next(i for i, l in enumerate(lines) if i + 1 >= start and pat in l) + 1
    bounds.append((ln, name))
def phase_of(line):
    if line < start or line >= bounds[-1][0]:
        return None
    ph = "setup"
    for ln, name in bounds:
        if line >= ln:
            ph = name
    return ph

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = "?"; hdr = None; cur = None; sass = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No":
        hdr = {n: i for i, n in enumerate(r)}; continue
    if hdr is None: continue
    if r[0]:
        cur = (fname, int(r[0])); continue
    if len(r) <= hdr["Instructions Executed"] or r[2] in ("...", "-", ""): continue
    try:
        addr = int(r[2], 16) if not r[2].isdigit() else int(r[2])
    except ValueError:
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[3])
    op = m.group(2) if m else "?"
    g = lambda k: int(r[hdr[k]] or 0) if k in hdr and r[hdr[k]] not in ("", "-") else 0
    st = {k[6:]: g(k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
    sass.append((addr, cur, op, g("Instructions Executed"), g("# Samples"), st))
sass.sort()
agg = collections.defaultdict(lambda: [0, 0, collections.Counter(), collections.Counter()])
ph = "other"
tot_i = tot_s = 0
for addr, (f, l), op, ni, ns, st in sass:
    if f.endswith("tdlo_taskq.cuh"):
        p = phase_of(l)
        ph = p if p else "other"
    a = agg[ph]
    a[0] += ni; a[1] += ns; a[2][op] += ni
    for k, v in st.items(): a[3][k] += v
    tot_i += ni; tot_s += ns
F64 = ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX")
print(f"total warp instructions {tot_i:.4g}, samples {tot_s}")
print(f"{'phase':>11s} {'inst%':>6s} {'smpl%':>6s} {'fp64%':>6s}  top ops / top stalls")
order = ["setup", "prefetch", "sphere", "scan", "neighbours", "window", "phaseA", "normalise", "phaseB", "epilogue", "other"]
for p in order:
    if p not in agg: continue
    a = agg[p]
    f64 = sum(a[2][o] for o in F64)
    ops = " ".join(f"{o}:{100*c/max(a[0],1):.0f}" for o, c in a[2].most_common(7))
    sts = " ".join(f"{o}:{100*c/max(a[1],1):.0f}" for o, c in a[3].most_common(5))
    print(f"{p:>11s} {100*a[0]/tot_i:6.2f} {100*a[1]/tot_s:6.2f} {100*f64/max(a[0],1):6.1f}  {ops}\n{'':>33s}{sts}")
tot = collections.Counter(); n = 0
for p in order:
    if p in agg and p not in ("other", "epilogue"):
        tot.update(agg[p][3]); n += agg[p][1]
print("E-step tile loop, all phases: samples", n, " ".join(f"{k}:{100*v/max(n,1):.1f}" for k, v in tot.most_common(14)))
