"""Static SASS instruction histogram of the E-step by phase (development aid; same phase markers as ncu_phase_map.py).
Usage: python scripts/sass_phase_static.py <variant .cu> [-DNAME...]"""
import collections, re, subprocess, sys, os
SRC = "trackdlo_b200/csrc/tdlo_taskq.cuh"
marks = [("prefetch", "software prefetch"), ("sphere", "---- nearest node"), ("scan", "double best = 1e300"),
         ("neighbours", "whole column underflows"), ("window", "---- node window"), ("phaseA", "double w = 0.0;"),
         ("normalise", "if (quirk) {"), ("phaseB", "---- phase B"), ("epilogue", "per-warp tile-loop cycles"),
         ("END", "// Visibility pre-pass over one chunk")]
lines = open(SRC).read().splitlines()
start = next(i for i, l in enumerate(lines) if "static __device__ void tq_estep_chunk" in l) + 1
bounds = [(next(i for i, l in enumerate(lines) if i + 1 >= start and pat in l) + 1, name) for name, pat in marks]
def phase_of(line):
    if line < start or line >= bounds[-1][0]: return None
    ph = "setup"
    for ln, name in bounds:
        if line >= ln: ph = name
    return ph
cu = sys.argv[1]
subprocess.check_call(["nvcc", "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-cubin", "-o", "/tmp/w/s.cubin", cu] + sys.argv[2:])
out = subprocess.run(["nvdisasm", "--print-line-info", "/tmp/w/s.cubin"], capture_output=True, text=True).stdout
agg = collections.defaultdict(collections.Counter); ph = "other"; cur = None
for l in out.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if m.group(1).endswith("tdlo_taskq.cuh"):
            p = phase_of(int(m.group(2))); ph = p if p else "other"
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
    if m: agg[ph][m.group(2).split(".")[0]] += 1
F64 = ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX")
tot = 0
for p in ["setup", "prefetch", "sphere", "scan", "neighbours", "window", "phaseA", "normalise", "phaseB", "epilogue"]:
    c = agg[p]; n = sum(c.values()); tot += n
    print(f"{p:>11s} {n:5d}  fp64 {sum(c[o] for o in F64):4d}  " + " ".join(f"{o}:{v}" for o, v in c.most_common(12)))
print("E-step static total", tot)
