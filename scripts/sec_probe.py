import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from trackdlo_b200 import api, synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 64
wl = synth.make_batch(F, n_nodes=50, n_points=20000)
dev = torch.device("cuda:0")
ctx = api.Context(max_frames=F, max_nodes=50, max_points_total=int(wl["x_offsets"][-1]))
d = {k: torch.from_numpy(np.ascontiguousarray(wl[k])).to(dev) for k in ("X", "x_offsets", "Y")}
Y0 = d["Y"].clone(); s2 = torch.zeros(F, dtype=torch.float64, device=dev)
it = torch.zeros(F, dtype=torch.int32, device=dev); st = torch.zeros(F, dtype=torch.int32, device=dev)
cb = api.CpdBatchC(F, 50, d["X"].data_ptr(), d["x_offsets"].data_ptr(), None, d["Y"].data_ptr(), s2.data_ptr(), None, None, None, None, None, it.data_ptr(), st.data_ptr())
cp = api.CpdParams(max_iter=50, tol=0.0).to_c()
stream = torch.cuda.current_stream()
def run():
    d["Y"].copy_(Y0); s2.zero_()
    ctx.cpd_lle_batched_raw(cb, cp, device=True, stream=stream.cuda_stream); torch.cuda.synchronize()
run(); ctx.profile_phases(True); run(); ph = ctx.profile_phases(False)
v = list(ph["rank0"].values()) + list(ph["others"].values())
names = ["search range", "scan", "neighbours", "window", "phase A", "normalise+wb", "phase B", "reduce+owner"]
tiles = v[11]
print("tiles timed", tiles)
tot = sum(v[:8])
for n, x in zip(names, v[:8]): print(f"  {n:14s} {x/tiles:8.0f} cycles/tile  {100*x/tot:5.1f}%")
print("  total", tot / tiles)
