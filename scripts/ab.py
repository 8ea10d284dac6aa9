"""Quick device-resident timing of a few shapes with optional context options: development A/B harness.
Usage (GPU box): python scripts/ab.py [shape ...] [key=value ...]     shapes: C2 C2x8 C4 C1 C3 C5"""
import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np, torch
import bench
from trackdlo_b200 import api
if os.environ.get("TDLO_AB_LIB"):          # development A/B: time another build of the library in the same gpurun call
    api.LIB_PATH = os.path.abspath(os.environ["TDLO_AB_LIB"])
dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
opts = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
shapes = [a for a in sys.argv[1:] if "=" not in a] or ["C2", "C2x8"]
def run(name, frames, nodes, points, tp, steps=5, distinct=None, occlusion=0.0):
    wl = bench.make_workload(0, frames, distinct=distinct, n_nodes=nodes, n_points=points, occlusion=occlusion)
    ctx = api.Context(max_frames=frames, max_nodes=nodes, max_points_total=int(wl["x_offsets"][-1]))
    for k, v in opts.items(): ctx.set_option(k, float(v))
    db = bench.DeviceBatch(api, torch, dev, wl, nodes)
    stream = torch.cuda.current_stream(); tpc = tp.to_c(); evs = []
    call = lambda: ctx.tracking_step_batched_raw(db.batch, tpc, device=True, stream=stream.cuda_stream)
    for s in range(steps + 2):
        db.reset(); flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); call(); e1.record(stream)
        if s > 1: evs.append((e0, e1))
    ctx.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in evs)
    ctx.profile_phases(True); db.reset(); call(); ph = ctx.profile_phases(False)
    iters = int(db.iters.sum()); c = ph["cycles"]; tot = sum(c.values()); n = ph["counts"]
    print(f"{name:6s} {os.environ.get('TDLO_DEV_SMEM_PAD', '')} {opts} launch={ctx.launch_info()} median {t[len(t)//2]:8.3f} ms min {t[0]:8.3f}  iters {iters}  cyc/tile {n['warp_tile_loop_cycles']/max(n['tiles'],1):7.0f} | " +
          " ".join(f"{k}={v/tot:.3f}" for k, v in c.items() if v / tot > 0.02), flush=True)
    if frames == 1:
        print("        thread-0 cycles per EM iteration: " + " ".join(f"{k}={v/iters:.0f}" for k, v in c.items() if k != "queue_wait"), flush=True)
    ctx.close()
for s in shapes:
    if s == "C2": run("C2", 64, 50, 20000, api.TrackParams(max_iter=50, tol=0.0))
    if s == "C2x8": run("C2x8", 512, 50, 20000, api.TrackParams(max_iter=50, tol=0.0), distinct=64, steps=3)
    if s == "C4": run("C4", 512, 50, 20000, api.TrackParams(), distinct=64, steps=3)
    if s == "C2x1": run("C2x1", 1, 50, 20000, api.TrackParams(max_iter=50, tol=0.0))
    if s == "C2x16": run("C2x16", 16, 50, 20000, api.TrackParams(max_iter=50, tol=0.0))
    if s == "P300": run("P300", 1, 45, 300, api.TrackParams())
    if s == "P1000": run("P1000", 1, 45, 1000, api.TrackParams())
    if s == "P3000": run("P3000", 1, 45, 3000, api.TrackParams())
    if s == "C1": run("C1", 1, 30, 2000, api.TrackParams(max_iter=20, tol=0.0))
    if s == "C3": run("C3", 1, 50, 50000, api.TrackParams(max_iter=50, tol=0.0), occlusion=0.4)
    if s == "C5": run("C5", 8, 200, 100000, api.TrackParams(max_iter=50, tol=0.0), distinct=2, steps=2)
