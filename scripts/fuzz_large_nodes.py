import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np
import oracle
from trackdlo_b200 import api, synth
def rel(a, b): return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))
ctx = api.Context(max_frames=3, max_nodes=256, max_points_total=3 * 6000)
worst = 0
for Nn in (65, 67, 96, 127, 129, 150, 199, 255, 256):
    for lle in (False, True):
        Mp = 2500 + 13 * Nn
        f = synth.make_frame(Nn, n_nodes=Nn, n_points=Mp)
        kw = dict(max_iter=3, tol=0.0, include_lle=lle)
        if lle: kw.update(beta=3.0, lambda_=1.0)
        o = oracle.cpd_lle(f["X"], f["Y"], 0.0, oracle.CpdParams(**kw))
        r = ctx.cpd_lle_batched(f["X"], np.array([0, Mp], np.int64), f["Y"][None], np.zeros(1), api.CpdParams(**kw))
        e, ew = rel(r["Y"][0], o["Y"]), rel(r["W"][0], o["W"])
        worst = max(worst, e)
        print(f"Nn={Nn:3d} lle={lle!s:5s} iters {r['iters'][0]} status {r['status'][0]} rel err Y {e:.2e} W {ew:.2e}" + ("" if e < 1e-6 and ew < 1e-5 else "  <-- CHECK"), flush=True)
# ragged batch with mixed node counts (one launch, NPASS=8 kernel)
frames = [synth.make_frame(5 + i, n_nodes=n, n_points=m) for i, (n, m) in enumerate(((70, 3000), (200, 5000), (33, 1000)))]
xo = np.zeros(4, np.int64); xo[1:] = np.cumsum([len(f["X"]) for f in frames])
Y = np.zeros((3, 256, 3)); nn = np.array([f["Y"].shape[0] for f in frames], np.int32)
for i, f in enumerate(frames): Y[i, :nn[i]] = f["Y"]
r = ctx.cpd_lle_batched(np.concatenate([f["X"] for f in frames]), xo, Y, np.zeros(3), api.CpdParams(max_iter=4, tol=0.0), n_nodes=nn)
for i, f in enumerate(frames):
    o = oracle.cpd_lle(f["X"], f["Y"], 0.0, oracle.CpdParams(max_iter=4, tol=0.0))
    e = rel(r["Y"][i, :nn[i]], o["Y"]); worst = max(worst, e)
    print(f"ragged frame {i} Nn={nn[i]} rel err Y {e:.2e}" + ("" if e < 1e-6 else "  <-- CHECK"))
print("worst", worst)
