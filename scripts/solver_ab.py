"""A/B timing of the M-step solver options (TDLO_OPT_SOLVER) on the BASELINE shapes, device-resident, with the kernel's
phase counters: cycles per M-step of the solve / gather+assemble / update phases.  Usage (GPU box): python scripts/solver_ab.py"""
import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np, torch
import bench
from trackdlo_b200 import api
dev = torch.device("cuda:0")
flush = torch.empty(256*1024*1024, dtype=torch.uint8, device=dev)
def run(name, frames, nodes, points, tp, solver, steps=4, distinct=None, occlusion=0.0, cpd=False):
    wl = bench.make_workload(0, frames, distinct=distinct, n_nodes=nodes, n_points=points, occlusion=occlusion)
    ctx = api.Context(max_frames=frames, max_nodes=nodes, max_points_total=int(wl["x_offsets"][-1]))
    ctx.set_option("solver", solver)
    db = bench.DeviceBatch(api, torch, dev, wl, nodes)
    stream = torch.cuda.current_stream(); tpc = tp.to_c(); evs=[]
    if cpd:
        it1 = torch.zeros(frames, dtype=torch.int32, device=dev)
        cb = api.CpdBatchC(frames, nodes, db.d["X"].data_ptr(), db.d["x_offsets"].data_ptr(), None, db.d["Y"].data_ptr(), db.s2.data_ptr(), None, None, None, None, None, it1.data_ptr(), db.status.data_ptr())
        cp = api.CpdParams(max_iter=tp.max_iter, tol=tp.tol).to_c()
        call = lambda: ctx.cpd_lle_batched_raw(cb, cp, device=True, stream=stream.cuda_stream)
    else:
        call = lambda: ctx.tracking_step_batched_raw(db.batch, tpc, device=True, stream=stream.cuda_stream)
    for s in range(steps+1):
        db.reset(); flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); call(); e1.record(stream)
        if s>0: evs.append((e0,e1))
    ctx.synchronize()
    ms = sum(a.elapsed_time(b) for a,b in evs)/steps
    ctx.profile_phases(True); db.reset(); call(); ph = ctx.profile_phases(False)
    iters = int(it1.sum()) if cpd else int(db.iters.sum())
    c = ph["cycles"]; tot = sum(c.values())
    print(f"{name:14s} solver={solver} {ms:8.3f} ms  iters {iters}  per M-step cycles: gather+assemble {c['mstep_gather_assemble']/iters:8.0f} solve {c['solve']/iters:8.0f} update {c['update']/iters:7.0f} | shares: " +
          " ".join(f"{k}={v/tot:.3f}" for k,v in c.items() if v/tot>0.02), flush=True)
    ctx.close()
which = sys.argv[1:] or ["C2", "C1", "C4", "C3", "C5", "C5cpd"]
for solver in (1, 0, 2):
    if "C2" in which: run("C2", 64, 50, 20000, api.TrackParams(max_iter=50, tol=0.0), solver)
    if "C1" in which: run("C1", 1, 30, 2000, api.TrackParams(max_iter=20, tol=0.0), solver)
    if "C4" in which: run("C4_shard", 512, 50, 20000, api.TrackParams(), solver, distinct=64, steps=3)
    if "C3" in which: run("C3", 1, 50, 50000, api.TrackParams(max_iter=50, tol=0.0), solver, occlusion=0.4)
    if "C5" in which: run("C5_shard", 8, 200, 100000, api.TrackParams(max_iter=50, tol=0.0), solver, distinct=2, steps=2)
    if "C5cpd" in which: run("C5_cpd_only", 8, 200, 100000, api.TrackParams(max_iter=50, tol=0.0), solver, distinct=2, steps=2, cpd=True)
