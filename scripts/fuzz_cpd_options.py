"""Randomised differential run of cpd_lle with correspondence priors, visibility weighting (k_vis), supplied H and a
given sigma2, ragged node counts per frame, against the oracle (development aid)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import oracle
from trackdlo_b200 import api, synth

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 11)


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))


S = 64
ctx = api.Context(max_frames=4, max_nodes=S, max_points_total=4 * 6000)
bad = 0; worst = 0.0
for case in range(n_cases):
    F = int(rng.integers(1, 5))
    ctx.set_option("chunk_points", int(rng.choice([0, 256, 1024])))
    lle = bool(rng.random() < 0.3)
    kw = dict(max_iter=int(rng.integers(1, 16)), tol=float(rng.choice([0.0, 2e-4])), include_lle=lle, alpha=3.0, k_vis=float(rng.choice([0.0, 50.0])),
              visibility_threshold=0.008)
    if lle:
        kw.update(beta=3.0, lambda_=1.0)
    frames, pri, npri, nvis, s2, Hs = [], np.zeros((F, S, 4)), np.zeros(F, np.int32), np.zeros(F, np.int32), np.zeros(F), np.zeros((F, S, S))
    useH = lle and rng.random() < 0.5
    for i in range(F):
        Nn = int(rng.integers(6, S + 1))
        f = synth.make_frame(int(rng.integers(0, 100000)), n_nodes=Nn, n_points=int(rng.integers(200, 6000)), occlusion=float(rng.choice([0.0, 0.3])))
        frames.append(f)
        k = int(rng.integers(0, Nn + 1)) if rng.random() < 0.6 else 0
        idx = np.sort(rng.choice(Nn, size=k, replace=False))
        pri[i, :k, 0] = idx; pri[i, :k, 1:] = f["Y"][idx] + rng.normal(0, 0.003, (k, 3)); npri[i] = k
        nvis[i] = int(rng.choice([0, Nn, max(1, Nn // 2), len(f["vis"])]))
        s2[i] = 10.0 ** rng.uniform(-6, -3) if rng.random() < 0.3 else 0.0
        if useH:
            Hs[i, :Nn, :Nn] = oracle.lle_H(f["Y"])
    xo = np.zeros(F + 1, np.int64); xo[1:] = np.cumsum([len(f["X"]) for f in frames])
    Y = np.zeros((F, S, 3)); nn = np.array([f["Y"].shape[0] for f in frames], np.int32)
    for i, f in enumerate(frames):
        Y[i, :nn[i]] = f["Y"]
    r = ctx.cpd_lle_batched(np.concatenate([f["X"] for f in frames]), xo, Y, s2, api.CpdParams(**kw), n_nodes=nn, priors=pri, n_priors=npri, n_visible=nvis,
                            H=Hs if useH else None)
    for i, f in enumerate(frames):
        o = oracle.cpd_lle(f["X"], f["Y"], float(s2[i]), oracle.CpdParams(**kw), priors=pri[i, :npri[i]], vis=np.arange(nvis[i]),
                           H=Hs[i, :nn[i], :nn[i]] if useH else None)
        n = nn[i]
        ok = r["iters"][i] == o["iters"] and bool(r["status"][i] & 1) == (not o["converged"])
        e = rel(r["Y"][i, :n], o["Y"]); worst = max(worst, e)
        if not ok or e > 1e-6:
            bad += 1
            print(f"case {case} frame {i}: Nn={n} Mp={len(f['X'])} lle={lle} H={useH} npri={npri[i]} nvis={nvis[i]} k_vis={kw['k_vis']} s2={s2[i]:.1e} it={kw['max_iter']}: "
                  f"iters {r['iters'][i]} vs {o['iters']} status {r['status'][i]} rel err {e:.2e}  <-- CHECK", flush=True)
print(f"{n_cases} batches, {bad} mismatching frames, worst rel err {worst:.2e}")
ctx.close()
