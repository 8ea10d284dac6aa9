import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np, torch
import bench
from trackdlo_b200 import api
dev = torch.device("cuda:0")
def run(nodes, points, solver, frames=1):
    wl = bench.make_workload(0, frames, distinct=1, n_nodes=nodes, n_points=points)
    ctx = api.Context(max_frames=frames, max_nodes=nodes, max_points_total=int(wl["x_offsets"][-1]))
    ctx.set_option("solver", solver)
    db = bench.DeviceBatch(api, torch, dev, wl, nodes)
    stream = torch.cuda.current_stream(); tpc = api.TrackParams(max_iter=10, tol=0.0).to_c()
    for _ in range(2):
        db.reset(); ctx.tracking_step_batched_raw(db.batch, tpc, device=True, stream=stream.cuda_stream)
    ctx.synchronize()
    ctx.profile_phases(True); db.reset(); ctx.tracking_step_batched_raw(db.batch, tpc, device=True, stream=stream.cuda_stream); ph = ctx.profile_phases(False)
    c = ph["cycles"]; it = int(db.iters.sum())
    print(f"Nn={nodes} Mp={points} frames={frames} solver={solver}: iters {it} per M-step cycles gather {c['mstep_gather_assemble']/it:.0f} solve {c['solve']/it:.0f} update {c['update']/it:.0f} start_call {c['start_call']/ (2*frames):.0f}", flush=True)
    ctx.close()
for nodes in (200, 100, 50):
    for solver in (2, 1):
        run(nodes, 20000, solver)
