"""Aggregates the ncu source page (ncu -i X.ncu-rep --page source --csv --print-source sass,cuda) by CUDA source line:
executed warp instructions, stall samples; prints the hottest lines.  Usage: python scripts/ncu_source_hot.py X.ncu-rep [N]"""
import csv, subprocess, sys, collections, re, io
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
lines = out.splitlines()
# find header row
hi = next(i for i, l in enumerate(lines) if l.startswith('"Line No"') or l.startswith('"Address"') or '"Instructions Executed"' in l)
rows = list(csv.reader(lines[hi:]))
hdr = rows[0]
col = {n: i for i, n in enumerate(hdr)}
# there are two "Source" columns: first is cuda source (line), second sass
src_idx = [i for i, n in enumerate(hdr) if n == "Source"]
ie = col["Instructions Executed"]; ns = col["# Samples"]; ln = col.get("Line No", 0)
agg = collections.OrderedDict()
tot_i = tot_s = 0
fp64_i = 0
per_line_fp64 = collections.Counter()
for r in rows[1:]:
    if len(r) <= ie: continue
    try: i = int(r[ie] or 0); s = int(r[ns] or 0)
    except ValueError: continue
    key = (r[ln], r[src_idx[0]].strip()[:110])
    a = agg.setdefault(key, [0, 0, collections.Counter()])
    a[0] += i; a[1] += s
    sass = r[src_idx[1]] if len(src_idx) > 1 else ""
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", sass)
    op = m.group(2).split(".")[0] if m else "?"
    a[2][op] += i
    tot_i += i; tot_s += s
    if op in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX", "MUFU", "F2F", "I2F", "F2I", "D2I", "I2D"): per_line_fp64[key] += i
    if op in ("DFMA", "DADD", "DMUL", "DSETP"): fp64_i += i
print(f"total warp instructions {tot_i:.4g}, samples {tot_s}, FP64-pipe (DFMA/DADD/DMUL/DSETP) {fp64_i:.4g} = {100*fp64_i/max(tot_i,1):.1f}%")
print(f"{'line':>5s} {'inst%':>6s} {'smpl%':>6s}  top ops | source")
for (l, s), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:N]:
    ops = " ".join(f"{o}:{100*c/max(a[0],1):.0f}" for o, c in a[2].most_common(4))
    print(f"{l:>5s} {100*a[0]/max(tot_i,1):6.2f} {100*a[1]/max(tot_s,1):6.2f}  {ops:40s} | {s}")
