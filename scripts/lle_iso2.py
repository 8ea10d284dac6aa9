import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np, torch
import bench
from trackdlo_b200 import api
dev = torch.device("cuda:0")
for nodes in (200, 100):
    wl = bench.make_workload(0, 1, distinct=1, n_nodes=nodes, n_points=5000)
    ctx = api.Context(max_frames=1, max_nodes=nodes, max_points_total=int(wl["x_offsets"][-1]))
    db = bench.DeviceBatch(api, torch, dev, wl, nodes)
    stream = torch.cuda.current_stream(); tpc = api.TrackParams(max_iter=10, tol=0.0).to_c()
    for _ in range(2):
        db.reset(); ctx.tracking_step_batched_raw(db.batch, tpc, device=True, stream=stream.cuda_stream)
    ctx.synchronize()
    ctx.profile_phases(True); db.reset(); ctx.tracking_step_batched_raw(db.batch, tpc, device=True, stream=stream.cuda_stream); ph = ctx.profile_phases(False)
    c = ph["counts"]; print(nodes, "LLE solves: 10; prologue", c["row_blocks"]/10, "forward", c["warp_tile_loop_cycles"]/10, "backward", c["end_of_task_reduction_cycles"]/10, "solve total per M-step avg", ph["cycles"]["solve"]/20)
    ctx.close()
