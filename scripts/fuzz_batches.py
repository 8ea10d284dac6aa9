"""Randomised differential run on BATCHES: several ragged frames per call (tracking_step), random engine options, random
sigma2_in, against the per-frame oracle (development aid).  Usage: python scripts/fuzz_batches.py [n_cases] [seed]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import oracle
from trackdlo_b200 import api, synth

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 3)


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))


ctx = api.Context(max_frames=6, max_nodes=64, max_points_total=6 * 7000)
worst = 0.0; bad = 0
for case in range(n_cases):
    F = int(rng.integers(1, 7)); Nn = int(rng.integers(8, 65))
    frames = []
    while len(frames) < F:
        f = synth.make_frame(int(rng.integers(0, 100000)), n_nodes=Nn, n_points=int(rng.integers(100, 7000)), occlusion=float(rng.choice([0.0, 0.15, 0.4])))
        if len(f["vis_ext"]) >= 4:
            frames.append(f)
    engine = "tq"
    ctx.set_option("chunk_points", int(rng.choice([0, 256, 1024, 2048])))
    ctx.set_option("threads", int(rng.choice([256, 256])))
    s2 = np.where(rng.random(F) < 0.3, 10.0 ** rng.uniform(-6, -3, F), 0.0)
    mi = int(rng.integers(1, 20)); tol = float(rng.choice([0.0, 2e-4]))
    xo = np.zeros(F + 1, np.int64); xo[1:] = np.cumsum([len(f["X"]) for f in frames])
    vo = np.zeros(F + 1, np.int64); vo[1:] = np.cumsum([len(f["vis"]) for f in frames])
    eo = np.zeros(F + 1, np.int64); eo[1:] = np.cumsum([len(f["vis_ext"]) for f in frames])
    r = ctx.tracking_step_batched(np.concatenate([f["X"] for f in frames]), xo, np.stack([f["Y"] for f in frames]), s2, np.stack([f["rest"] for f in frames]),
                                  np.concatenate([f["vis"] for f in frames]), vo, np.concatenate([f["vis_ext"] for f in frames]), eo,
                                  api.TrackParams(max_iter=mi, tol=tol))
    for i, f in enumerate(frames):
        o = oracle.tracking_step(f["X"], f["Y"], float(s2[i]), f["rest"], f["vis"], f["vis_ext"], oracle.TrackParams(max_iter=mi, tol=tol))
        ok = list(r["iters"][i]) == list(o["iters"]) and r["state"][i] == o["state"] and r["n_priors"][i] == len(o["priors"])
        e = rel(r["Y"][i], o["Y"]); es = abs(r["sigma2"][i] - o["sigma2"]) / o["sigma2"]
        worst = max(worst, e)
        if not ok or e > 1e-6 or es > 1e-5:
            bad += 1
            print(f"case {case} frame {i}/{F} engine {engine} Nn={Nn} Mp={len(f['X'])} s2in={s2[i]:.2e} it={mi} tol={tol:g}: iters {list(r['iters'][i])} vs {list(o['iters'])} "
                  f"state {r['state'][i]} vs {o['state']} status {r['status'][i]} rel err {e:.2e} sigma2 err {es:.2e}  <-- CHECK", flush=True)
print(f"{n_cases} batches, {bad} mismatching frames, worst rel err {worst:.2e}")
ctx.close()
