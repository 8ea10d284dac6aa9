import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np, torch
from trackdlo_b200 import api, synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 256
wl = synth.make_batch(F, n_nodes=50, n_points=20000)
dev = torch.device("cuda:0")
ctx = api.Context(max_frames=F, max_nodes=50, max_points_total=int(wl["x_offsets"][-1]))
d = {k: torch.from_numpy(np.ascontiguousarray(wl[k])).to(dev) for k in ("X", "x_offsets", "Y", "rest", "vis", "vis_offsets", "vis_ext", "vis_ext_offsets")}
Y0 = d["Y"].clone(); s2 = torch.zeros(F, dtype=torch.float64, device=dev)
it = torch.zeros(F, 2, dtype=torch.int32, device=dev); st = torch.zeros(F, dtype=torch.int32, device=dev)
tb = api.TrackBatchC(F, 50, d["X"].data_ptr(), d["x_offsets"].data_ptr(), d["Y"].data_ptr(), s2.data_ptr(), d["rest"].data_ptr(),
                     d["vis"].data_ptr(), d["vis_offsets"].data_ptr(), d["vis_ext"].data_ptr(), d["vis_ext_offsets"].data_ptr(),
                     None, None, None, None, it.data_ptr(), st.data_ptr(), None)
stream = torch.cuda.current_stream()
def run(tp):
    best = 1e9
    for r in range(3):
        d["Y"].copy_(Y0); s2.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); ctx.tracking_step_batched_raw(tb, tp, device=True, stream=stream.cuda_stream); e1.record(stream); torch.cuda.synchronize()
        if r: best = min(best, e0.elapsed_time(e1))
    return best
for name, tp in (("default tol", api.TrackParams().to_c()), ("fixed 50+50", api.TrackParams(max_iter=50, tol=0.0).to_c())):
    for infl in (0,):
        for chunk in (0, 1024, 2048, 4096):
            ctx.set_option("inflight", infl); ctx.set_option("chunk_points", chunk)
            ms = run(tp)
            print(f"{name:12s} F={F} inflight={infl:3d} chunk={chunk}: {ms:7.2f} ms  {F/ms*1e3:8.0f} frames/s  {int(it.sum())/ms*1e3:9.0f} it/s", flush=True)
