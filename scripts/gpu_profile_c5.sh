#!/bin/bash
# ncu --set full capture of the Nn=200 dense-solve stress shape (C5 shard, reduced: 2 frames, 10 iterations)
mkdir -p gpurun_out
cat > /tmp/c5_run.py <<'PY'
import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np, torch
from trackdlo_b200 import api, synth
F = 2
wl = synth.make_batch(F, n_nodes=200, n_points=100000)
dev = torch.device("cuda:0")
ctx = api.Context(max_frames=F, max_nodes=200, max_points_total=int(wl["x_offsets"][-1]))
d = {k: torch.from_numpy(np.ascontiguousarray(wl[k])).to(dev) for k in ("X", "x_offsets", "Y")}
s2 = torch.zeros(F, dtype=torch.float64, device=dev); it = torch.zeros(F, dtype=torch.int32, device=dev); st = torch.zeros(F, dtype=torch.int32, device=dev)
cb = api.CpdBatchC(F, 200, d["X"].data_ptr(), d["x_offsets"].data_ptr(), None, d["Y"].data_ptr(), s2.data_ptr(), None, None, None, None, None, it.data_ptr(), st.data_ptr())
for _ in range(2):
    ctx.cpd_lle_batched_raw(cb, api.CpdParams(max_iter=10, tol=0.0).to_c(), device=True, stream=torch.cuda.current_stream().cuda_stream); torch.cuda.synchronize()
print("iters", it.cpu().numpy())
PY
ncu --set full --clock-control none --import-source on -k regex:tdlo_tq -s 1 -c 1 -o gpurun_out/prof_c5 -f python /tmp/c5_run.py > gpurun_out/prof_c5.log 2>&1
tail -2 gpurun_out/prof_c5.log
