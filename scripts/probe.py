"""GPU probe: CUDA path vs the CPU oracle on a few seeded frames (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from trackdlo_b200 import api, synth


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def check_cpd(ctx, frames, prm_o, prm_g, tag, priors=None, nvis=None, cluster=0):
    ctx.set_cluster_size(cluster)
    F = len(frames)
    X = np.concatenate([f["X"] for f in frames]); xo = np.zeros(F + 1, np.int64)
    xo[1:] = np.cumsum([f["X"].shape[0] for f in frames])
    Y = np.stack([f["Y"] for f in frames])
    S = Y.shape[1]
    pr = npr = None
    if priors is not None:
        pr = np.zeros((F, S, 4)); npr = np.zeros(F, np.int32)
        for i, p in enumerate(priors):
            pr[i, :len(p)] = p; npr[i] = len(p)
    t = time.time()
    g = ctx.cpd_lle_batched(X, xo, Y, np.zeros(F), prm_g, priors=pr, n_priors=npr, n_visible=nvis)
    tg = time.time() - t
    worst = 0
    for i, f in enumerate(frames):
        o = oracle.cpd_lle(f["X"], f["Y"], 0.0, prm_o, priors=None if priors is None else priors[i],
                           vis=None if nvis is None else np.arange(nvis[i]))
        eY, eW, eS = rel(g["Y"][i], o["Y"]), rel(g["W"][i], o["W"]), abs(g["sigma2"][i] - o["sigma2"]) / o["sigma2"]
        worst = max(worst, eY, eW, eS)
        print(f"  [{tag}] frame {i}: iters gpu/cpu {g['iters'][i]}/{o['iters']} status {g['status'][i]} conv {o['converged']}"
              f" relY {eY:.2e} relW {eW:.2e} relS2 {eS:.2e}")
    print(f"[{tag}] cluster={ctx.launch_info()} time {tg*1e3:.1f} ms worst {worst:.2e}")
    return worst


def main():
    ctx = api.Context(max_frames=64, max_nodes=64, max_points_total=64 * 20000)
    print(api.load_library().tdlo_version().decode())
    fr = [synth.make_frame(i, n_nodes=30, n_points=2000) for i in range(2)]
    po = oracle.CpdParams(max_iter=20, tol=0.0); pg = api.CpdParams(max_iter=20, tol=0.0)
    for c in (1, 2, 4):
        check_cpd(ctx, fr, po, pg, f"C1 fixed20 c{c}", cluster=c)
    po = oracle.CpdParams(); pg = api.CpdParams()
    check_cpd(ctx, fr, po, pg, "C1 converge")
    # LLE path
    po = oracle.CpdParams(beta=3.0, lambda_=1.0, include_lle=True, max_iter=20, tol=0.0)
    pg = api.CpdParams(beta=3.0, lambda_=1.0, include_lle=True, max_iter=20, tol=0.0)
    check_cpd(ctx, fr, po, pg, "C1 lle")
    # vis branch + priors
    fo = [synth.make_frame(i, n_nodes=50, n_points=6000, occlusion=0.4) for i in range(2)]
    po = oracle.CpdParams(alpha=3.0, k_vis=50.0, visibility_threshold=0.008, max_iter=20, tol=0.0)
    pg = api.CpdParams(alpha=3.0, k_vis=50.0, visibility_threshold=0.008, max_iter=20, tol=0.0)
    pri = [np.concatenate([np.arange(0, 50, 7)[:, None].astype(float), f["Y"][::7] + 0.002], axis=1) for f in fo]
    nv = np.array([len(f["vis_ext"]) for f in fo], np.int32)
    check_cpd(ctx, fo, po, pg, "vis+priors", priors=pri, nvis=nv)
    # tracking step
    tp_o = oracle.TrackParams(); tp_g = api.TrackParams()
    for frames, tag in ((fr, "track C1"), (fo, "track occl")):
        b = dict(X=np.concatenate([f["X"] for f in frames]))
        F = len(frames)
        xo = np.zeros(F + 1, np.int64); xo[1:] = np.cumsum([f["X"].shape[0] for f in frames])
        vo = np.zeros(F + 1, np.int64); vo[1:] = np.cumsum([len(f["vis"]) for f in frames])
        eo = np.zeros(F + 1, np.int64); eo[1:] = np.cumsum([len(f["vis_ext"]) for f in frames])
        g = ctx.tracking_step_batched(b["X"], xo, np.stack([f["Y"] for f in frames]), np.zeros(F),
                                      np.stack([f["rest"] for f in frames]),
                                      np.concatenate([f["vis"] for f in frames]), vo,
                                      np.concatenate([f["vis_ext"] for f in frames]), eo, tp_g)
        for i, f in enumerate(frames):
            o = oracle.tracking_step(f["X"], f["Y"], 0.0, f["rest"], f["vis"], f["vis_ext"], tp_o)
            V = len(f["vis_ext"])
            print(f"  [{tag}] frame {i}: state {g['state'][i]}/{o['state']} iters {g['iters'][i]}/{o['iters']} status {g['status'][i]} err {o['err']}"
                  f" npri {g['n_priors'][i]}/{len(o['priors'])} relGuide {rel(g['guide'][i][:V], o['guide']):.2e}"
                  f" relPri {rel(g['priors'][i][:len(o['priors'])], o['priors']) if len(o['priors'])==g['n_priors'][i] else -1:.2e}"
                  f" relY {rel(g['Y'][i], o['Y']):.2e} relS2 {abs(g['sigma2'][i]-o['sigma2'])/o['sigma2']:.2e}")
    # C2-shaped timing
    F = 8
    frs = [synth.make_frame(i, n_nodes=50, n_points=20000) for i in range(F)]
    pg = api.CpdParams(max_iter=50, tol=0.0)
    X = np.concatenate([f["X"] for f in frs]); xo = np.zeros(F + 1, np.int64); xo[1:] = np.cumsum([f["X"].shape[0] for f in frs])
    Y = np.stack([f["Y"] for f in frs])
    ctx.set_cluster_size(0)
    for rep in range(3):
        t = time.time(); g = ctx.cpd_lle_batched(X, xo, Y, np.zeros(F), pg); dt = time.time() - t
        print(f"C2x{F}: {dt*1e3:.2f} ms  -> {F*50/dt:.0f} it/s  info {ctx.launch_info()} iters {g['iters'][:3]} status {g['status'][:3]}")
    t = time.time(); o = oracle.cpd_lle(frs[0]["X"], frs[0]["Y"], 0.0, oracle.CpdParams(max_iter=50, tol=0.0)); dt = time.time() - t
    print(f"oracle 1 frame 50 it: {dt:.2f}s relY {rel(g['Y'][0], o['Y']):.2e} relW {rel(g['W'][0], o['W']):.2e}")


if __name__ == "__main__":
    main()
