"""BASELINE.json configs C1..C5 on ONE GPU: time through the device-pointer C ABI (CUDA events, inputs resident) and,
where the CPU oracle finishes in seconds, parity against it.  Writes one JSON line per case.
C4 / C5 are the per-GPU shards of the 8-GPU configs (512 frames; 8 frames at Nn=200, Mp=100000).
Usage: python scripts/config_sweep.py [--quick]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import oracle
from trackdlo_b200 import api, synth

quick = "--quick" in sys.argv
dev = torch.device("cuda:0")


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))


def run_track(ctx, wl, F, N, tp, reps=3):
    d = {k: torch.from_numpy(np.ascontiguousarray(wl[k])).to(dev) for k in ("X", "x_offsets", "Y", "rest", "vis", "vis_offsets", "vis_ext", "vis_ext_offsets")}
    Y0 = d["Y"].clone(); s2 = torch.zeros(F, dtype=torch.float64, device=dev)
    it = torch.zeros(F, 2, dtype=torch.int32, device=dev); st = torch.zeros(F, dtype=torch.int32, device=dev)
    tb = api.TrackBatchC(F, N, d["X"].data_ptr(), d["x_offsets"].data_ptr(), d["Y"].data_ptr(), s2.data_ptr(), d["rest"].data_ptr(),
                         d["vis"].data_ptr(), d["vis_offsets"].data_ptr(), d["vis_ext"].data_ptr(), d["vis_ext_offsets"].data_ptr(),
                         None, None, None, None, it.data_ptr(), st.data_ptr(), None)
    stream = torch.cuda.current_stream(); best = 1e30
    for r in range(reps + 1):
        d["Y"].copy_(Y0); s2.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); ctx.tracking_step_batched_raw(tb, tp.to_c(), device=True, stream=stream.cuda_stream); e1.record(stream)
        torch.cuda.synchronize()
        if r: best = min(best, e0.elapsed_time(e1))
    return best, d["Y"].cpu().numpy(), s2.cpu().numpy(), it.cpu().numpy(), st.cpu().numpy()


def run_cpd(ctx, wl, F, N, cp, reps=3):
    d = {k: torch.from_numpy(np.ascontiguousarray(wl[k])).to(dev) for k in ("X", "x_offsets", "Y")}
    Y0 = d["Y"].clone(); s2 = torch.zeros(F, dtype=torch.float64, device=dev)
    it = torch.zeros(F, dtype=torch.int32, device=dev); st = torch.zeros(F, dtype=torch.int32, device=dev)
    W = torch.zeros(F, N, 3, dtype=torch.float64, device=dev)
    cb = api.CpdBatchC(F, N, d["X"].data_ptr(), d["x_offsets"].data_ptr(), None, d["Y"].data_ptr(), s2.data_ptr(), None, None, None, None,
                       W.data_ptr(), it.data_ptr(), st.data_ptr())
    stream = torch.cuda.current_stream(); best = 1e30
    for r in range(reps + 1):
        d["Y"].copy_(Y0); s2.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); ctx.cpd_lle_batched_raw(cb, cp.to_c(), device=True, stream=stream.cuda_stream); e1.record(stream)
        torch.cuda.synchronize()
        if r: best = min(best, e0.elapsed_time(e1))
    return best, d["Y"].cpu().numpy(), W.cpu().numpy(), s2.cpu().numpy(), it.cpu().numpy(), st.cpu().numpy()


def emit(**kw):
    print(json.dumps(kw), flush=True)


# ---- C1: single frame Nn=30 Mp=2000, 20 iterations (correctness gate)
wl = synth.make_batch(1, n_nodes=30, n_points=2000)
ctx = api.Context(max_frames=1, max_nodes=30, max_points_total=2000)
cp = api.CpdParams(max_iter=20, tol=0.0)
ms, Y, W, s2, it, st = run_cpd(ctx, wl, 1, 30, cp)
o = oracle.cpd_lle(wl["frames"][0]["X"], wl["frames"][0]["Y"], 0.0, oracle.CpdParams(max_iter=20, tol=0.0))
emit(config="C1 cpd_lle", nodes=30, points=2000, frames=1, iters=int(it.sum()), ms=ms, it_per_s=it.sum() / ms * 1e3,
     rel_err_Y=rel(Y[0], o["Y"]), rel_err_W=rel(W[0], o["W"]))
tp = api.TrackParams(max_iter=20, tol=0.0)
ms, Y, s2, it, st = run_track(ctx, wl, 1, 30, tp)
f = wl["frames"][0]
o = oracle.tracking_step(f["X"], f["Y"], 0.0, f["rest"], f["vis"], f["vis_ext"], oracle.TrackParams(max_iter=20, tol=0.0))
emit(config="C1 tracking_step", nodes=30, points=2000, frames=1, iters=int(it.sum()), ms=ms, it_per_s=it.sum() / ms * 1e3, rel_err_Y=rel(Y[0], o["Y"]))
ctx.close()

# ---- C3: Nn=50, Mp0=50000, 40 % occlusion (visibility branch + priors), fixed 50 iterations and converge at tol=2e-4
wl = synth.make_batch(1, n_nodes=50, n_points=50000, occlusion=0.4)
f = wl["frames"][0]
ctx = api.Context(max_frames=1, max_nodes=50, max_points_total=int(wl["x_offsets"][-1]))
for name, kw in (("fixed 50 it", dict(max_iter=50, tol=0.0)), ("tol=2e-4", dict(max_iter=50, tol=2e-4))):
    ms, Y, s2, it, st = run_track(ctx, wl, 1, 50, api.TrackParams(**kw))
    o = oracle.tracking_step(f["X"], f["Y"], 0.0, f["rest"], f["vis"], f["vis_ext"], oracle.TrackParams(**kw))
    emit(config="C3 tracking_step " + name, nodes=50, points=int(wl["x_offsets"][-1]), visible=len(f["vis"]), frames=1, iters=[int(v) for v in it[0]],
         oracle_iters=[int(v) for v in o["iters"]], ms=ms, it_per_s=it.sum() / ms * 1e3, rel_err_Y=rel(Y[0], o["Y"]), status=int(st[0]), state=int(o["state"]))
ctx.close()

# ---- C2 / C4 shard: 64 and 512 frames, Nn=50, Mp=20000
for F, kw, tag in ((64, dict(max_iter=50, tol=0.0), "C2 (64 frames, fixed 50+50 it)"),
                   (64 if quick else 512, dict(), "C4 shard (512 frames/GPU, default tol=2e-4, max_iter=50)")):
    t0 = time.time(); wl = synth.make_batch(F, n_nodes=50, n_points=20000); gen = time.time() - t0
    ctx = api.Context(max_frames=F, max_nodes=50, max_points_total=int(wl["x_offsets"][-1]))
    ms, Y, s2, it, st = run_track(ctx, wl, F, 50, api.TrackParams(**kw))
    emit(config=tag, nodes=50, points=20000, frames=F, iters=int(it.sum()), ms=ms, it_per_s=it.sum() / ms * 1e3, frames_per_s=F / ms * 1e3,
         status_or=int(np.bitwise_or.reduce(st)), gen_s=gen)
    ctx.close()

# ---- C5 shard: 8 frames, Nn=200, Mp=100000, 50 fixed iterations (dense-solve stress)
F = 2 if quick else 8
wl = synth.make_batch(F, n_nodes=200, n_points=100000)
ctx = api.Context(max_frames=F, max_nodes=200, max_points_total=int(wl["x_offsets"][-1]))
ms, Y, W, s2, it, st = run_cpd(ctx, wl, F, 200, api.CpdParams(max_iter=50, tol=0.0), reps=1)
emit(config="C5 shard cpd_lle (8 frames/GPU)", nodes=200, points=100000, frames=F, iters=int(it.sum()), ms=ms, it_per_s=it.sum() / ms * 1e3, status_or=int(np.bitwise_or.reduce(st)))
ms3, Y3, W3, s23, it3, st3 = run_cpd(ctx, dict(X=wl["frames"][0]["X"], x_offsets=np.array([0, len(wl["frames"][0]["X"])], np.int64), Y=wl["frames"][0]["Y"][None]), 1, 200,
                                     api.CpdParams(max_iter=3, tol=0.0), reps=1)
o = oracle.cpd_lle(wl["frames"][0]["X"], wl["frames"][0]["Y"], 0.0, oracle.CpdParams(max_iter=3, tol=0.0))
emit(config="C5 parity (1 frame, 3 iterations)", nodes=200, points=100000, ms=ms3, rel_err_Y=rel(Y3[0], o["Y"]), rel_err_W=rel(W3[0], o["W"]))
ctx.close()
