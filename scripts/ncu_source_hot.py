"""Hot source lines of a kernel from an .ncu-rep (needs -lineinfo and --import-source on).
Parses `ncu --page source --csv --print-source sass,cuda`: source-line rows followed by their SASS rows.
Usage: python scripts/ncu_source_hot.py X.ncu-rep [N]"""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 45
BY_SAMPLES = len(sys.argv) > 3 and sys.argv[3] == "samples"      # sort by stall samples instead of executed instructions
FILTER = sys.argv[4] if len(sys.argv) > 4 else ""                # only lines of files whose name contains this
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = "?"
hdr = None
agg = {}
ops_tot = collections.Counter()
tot_i = tot_s = 0
cur = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = {n: i for i, n in enumerate(r)}
        ie, ns = hdr["Instructions Executed"], hdr["# Samples"]
        continue
    if hdr is None or len(r) <= ie:
        continue
    if r[0]:                                  # a source line row
        cur = (fname, r[0], r[1].strip()[:100])
        agg.setdefault(cur, [0, 0, collections.Counter()])
        continue
    if r[2] in ("...", "-", ""):
        continue
    try:
        i = int(r[ie] or 0); s = int(r[ns] or 0)
    except ValueError:
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[3])
    op = m.group(2) if m else "?"
    a = agg[cur]
    a[0] += i; a[1] += s; a[2][op] += i
    ops_tot[op] += i
    tot_i += i; tot_s += s
print(f"total warp instructions {tot_i:.4g}, stall samples {tot_s}")
print("opcode mix: " + "  ".join(f"{o}:{100*c/tot_i:.1f}%" for o, c in ops_tot.most_common(22)))
f64 = sum(c for o, c in ops_tot.items() if o in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX"))
print(f"FP64-pipe instructions (DFMA/DADD/DMUL/DSETP/DMNMX): {f64:.4g} = {100*f64/tot_i:.1f}%")
print(f"{'file:line':>24s} {'inst%':>6s} {'smpl%':>6s}  top ops | source")
rows_ = [kv for kv in agg.items() if FILTER in kv[0][0]]
for (f, l, s), a in sorted(rows_, key=lambda kv: -kv[1][1 if BY_SAMPLES else 0])[:N]:
    ops = " ".join(f"{o}:{100*c/max(a[0],1):.0f}" for o, c in a[2].most_common(4))
    print(f"{f[-18:]+':'+l:>24s} {100*a[0]/max(tot_i,1):6.2f} {100*a[1]/max(tot_s,1):6.2f}  {ops:36s} | {s}")
