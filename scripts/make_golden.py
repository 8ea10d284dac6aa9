"""Mints tests/golden/*.npz: seeded inputs, the CPU oracle's outputs (oracle/trackdlo_oracle.cpp) and -- fields
`ref_*` -- the outputs of the REFERENCE ITSELF run here: oracle/_ref/libtrackdlo_ref.so is the unmodified
/root/reference/trackdlo/src/trackdlo.cpp + utils.cpp compiled by `make -C oracle _ref` (Eigen calls served by
oracle/ref_shim/eigen, or by the real Eigen with EIGEN_INCLUDE=...; the field `ref_eigen_shim` records which).

The reference ships no golden vectors of its own.  /root/reference does not exist on the GPU box, so these files are
how the reference's results travel there.  Regenerate with `make -C oracle all _ref && python scripts/make_golden.py`.
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from oracle import ref
from trackdlo_b200 import synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CPD_CASES = {
    # name: (frame kwargs, frame idx, CpdParams kwargs, priors?, n_visible?)
    "c1_fixed20": (dict(n_nodes=30, n_points=2000), 0, dict(max_iter=20, tol=0.0), False, False),
    "c1_converge": (dict(n_nodes=30, n_points=2000), 1, dict(), False, False),
    "c1_lle_preproc": (dict(n_nodes=30, n_points=2000), 0,
                       dict(beta=3.0, lambda_=1.0, include_lle=True, max_iter=20, tol=0.0), False, False),
    "occl_vis_priors": (dict(n_nodes=50, n_points=6000, occlusion=0.4), 3,
                        dict(alpha=3.0, k_vis=50.0, visibility_threshold=0.008, max_iter=20, tol=0.0), True, True),
    "n64_sigma_given": (dict(n_nodes=64, n_points=3000), 5, dict(max_iter=10, tol=0.0), False, False),
}
TRACK_CASES = {
    "track_c1": (dict(n_nodes=30, n_points=2000), 0),
    "track_c1_b": (dict(n_nodes=30, n_points=2000), 1),
    "track_occl_head": (dict(n_nodes=50, n_points=6000, occlusion=0.4), 0),
    "track_occl_mid": (dict(n_nodes=50, n_points=6000, occlusion=0.4), 1),
    "track_all_visible": (dict(n_nodes=40, n_points=8000, tau_vis=0.02), 2),
    # trackdlo.cpp:968-973 "Tail occluded" (state 2) and :980-995 "Both ends occluded" (state 4 -> traverse_euclidean
    # alignment 2, :749-895); _b: two visible runs; _lists: visible_nodes != visible_nodes_extended (:986-990)
    "track_occl_tail": (dict(n_nodes=50, n_points=6000, occl_windows=[(0.7, 1.0)]), 1),
    "track_occl_both": (dict(n_nodes=50, n_points=6000, occl_windows=[(0.0, 0.2), (0.8, 1.0)]), 0),
    "track_occl_both_b": (dict(n_nodes=50, n_points=6000, occl_windows=[(0.0, 0.15), (0.45, 0.6), (0.85, 1.0)]), 2),
    "track_occl_both_lists": (dict(n_nodes=50, n_points=5000, occl_windows=[(0.0, 0.2), (0.47, 0.53), (0.8, 1.0)]), 7),
}


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, (fkw, idx, pkw, use_pri, use_vis) in CPD_CASES.items():
        f = synth.make_frame(idx, **fkw)
        prm = oracle.CpdParams(**pkw)
        Nn = f["Y"].shape[0]
        priors = None
        if use_pri:
            sel = np.arange(0, Nn, 7)
            priors = np.concatenate([sel[:, None].astype(float), f["Y"][sel] + 0.002], axis=1)
        vis = f["vis_ext"] if use_vis else None
        s2_in = 0.0 if name != "n64_sigma_given" else 1e-4
        r = oracle.cpd_lle(f["X"], f["Y"], s2_in, prm, priors=priors, vis=vis, trace=True)
        rr = ref.cpd_lle(f["X"].astype(np.float32).astype(np.float64), f["Y"], s2_in, prm, priors=priors, vis=vis)
        np.savez_compressed(
            os.path.join(OUT, f"cpd_{name}.npz"),
            X=f["X"].astype(np.float32), Y_in=f["Y"], sigma2_in=s2_in,
            params=np.array([prm.beta, prm.lambda_, prm.lle_weight, prm.mu, prm.tol, prm.alpha, prm.k_vis,
                             prm.visibility_threshold, prm.max_iter, int(prm.include_lle)], float),
            priors=np.zeros((0, 4)) if priors is None else priors,
            n_visible=-1 if vis is None else len(vis),
            Y=r["Y"], W=r["W"], sigma2=r["sigma2"], iters=r["iters"], converged=int(r["converged"]), kept=r["kept"],
            tr_sigma2=r["trace"]["sigma2"], tr_Np=r["trace"]["Np"], tr_P1=r["trace"]["P1"], tr_PX=r["trace"]["PX"],
            A0=r["trace"]["A"], B0=r["trace"]["B"],
            ref_Y=rr["Y"], ref_sigma2=rr["sigma2"], ref_iters=rr["iters"], ref_converged=int(rr["converged"]),
            ref_eigen_shim=int(ref.uses_eigen_shim()))
        print(name, "iters", r["iters"], "converged", r["converged"], "sigma2", r["sigma2"], "kept", r["kept"])
    tp = oracle.TrackParams()
    for name, (fkw, idx) in TRACK_CASES.items():
        f = synth.make_frame(idx, **fkw)
        r = oracle.tracking_step(f["X"], f["Y"], 0.0, f["rest"], f["vis"], f["vis_ext"], tp)
        assert r["err"] == 0, name         # no input here may drive the reference into its out-of-range reads
        rr = ref.tracking_step(f["X"], f["Y"], 0.0, f["rest"], f["vis"], f["vis_ext"], tp)
        assert rr["state"] == r["state"] and list(rr["iters"]) == list(r["iters"]), name
        np.savez_compressed(
            os.path.join(OUT, f"{name}.npz"),
            X=f["X"].astype(np.float32), Y_in=f["Y"], rest=f["rest"], vis=f["vis"], vis_ext=f["vis_ext"],
            Y=r["Y"], sigma2=r["sigma2"], guide=r["guide"], priors=r["priors"], iters=r["iters"],
            converged=r["converged"], state=r["state"], err=r["err"],
            ref_Y=rr["Y"], ref_sigma2=rr["sigma2"], ref_guide=rr["guide"], ref_priors=rr["priors"], ref_iters=rr["iters"],
            ref_state=rr["state"], ref_eigen_shim=int(ref.uses_eigen_shim()))
        print(name, "state", r["state"], "iters", r["iters"], "err", r["err"], "npriors", len(r["priors"]))


if __name__ == "__main__":
    main()
