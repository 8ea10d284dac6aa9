import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np
import oracle
from trackdlo_b200 import api, synth
def rel(a, b): return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))
one = lambda n: np.array([0, n], np.int64)
for Nn, Mp, occ in ((100, 6000, 0.0), (100, 6000, 0.3), (150, 8000, 0.45), (72, 4000, 0.2)):
    f = synth.make_frame(Nn, n_nodes=Nn, n_points=Mp, occlusion=occ)
    ctx = api.Context(max_frames=1, max_nodes=Nn, max_points_total=len(f["X"]))
    for eng in ("tq",):
        tp = dict(max_iter=6, tol=0.0)
        r = ctx.tracking_step_batched(f["X"], one(len(f["X"])), f["Y"][None], np.zeros(1), f["rest"][None], f["vis"], one(len(f["vis"])), f["vis_ext"], one(len(f["vis_ext"])), api.TrackParams(**tp))
        o = oracle.tracking_step(f["X"], f["Y"], 0.0, f["rest"], f["vis"], f["vis_ext"], oracle.TrackParams(**tp))
        e = rel(r["Y"][0], o["Y"])
        print(f"Nn={Nn} occ={occ} V={len(f['vis_ext'])} engine={eng}: iters {list(r['iters'][0])} vs {list(o['iters'])} state {r['state'][0]}/{o['state']} status {r['status'][0]} rel err Y {e:.2e}" + ("" if e < 1e-5 else "  <-- CHECK"), flush=True)
    ctx.close()
