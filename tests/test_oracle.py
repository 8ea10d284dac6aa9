"""CPU tests of the oracle (oracle/trackdlo_oracle.cpp): against the committed goldens, against
the independent NumPy/LAPACK twin, and unit cases for the reference quirks listed in SURVEY.md §8c.
The reference itself has no tests/goldens (parity unpinned); see scripts/make_golden.py."""
import glob
import os

import numpy as np
import pytest

import oracle
from oracle import numpy_twin as nt
from trackdlo_b200 import synth


def _params_from(arr):
    return oracle.CpdParams(beta=arr[0], lambda_=arr[1], lle_weight=arr[2], mu=arr[3], tol=arr[4], alpha=arr[5],
                            k_vis=arr[6], visibility_threshold=arr[7], max_iter=int(arr[8]), include_lle=bool(arr[9]))


def _rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300))


def test_goldens_exist(golden_dir):
    assert len(glob.glob(os.path.join(golden_dir, "cpd_*.npz"))) >= 5
    assert len(glob.glob(os.path.join(golden_dir, "track_*.npz"))) >= 5


@pytest.mark.parametrize("name", ["c1_fixed20", "c1_converge", "c1_lle_preproc", "occl_vis_priors", "n64_sigma_given"])
def test_oracle_reproduces_cpd_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"cpd_{name}.npz"))
    prm = _params_from(g["params"])
    nvis = int(g["n_visible"])
    r = oracle.cpd_lle(g["X"].astype(np.float64), g["Y_in"], float(g["sigma2_in"]), prm,
                       priors=g["priors"] if len(g["priors"]) else None,
                       vis=np.arange(nvis) if nvis >= 0 else None, trace=True)
    assert r["iters"] == int(g["iters"]) and int(r["converged"]) == int(g["converged"]) and r["kept"] == int(g["kept"])
    # same source, possibly another libm / compiler: allow a few ulps amplified by the solve
    tol = 1e-6 if prm.include_lle else 1e-9      # LLE weights are rounding-noise driven
    assert _rel(r["Y"], g["Y"]) < tol
    assert _rel(r["W"], g["W"]) < tol * 100
    assert abs(r["sigma2"] - float(g["sigma2"])) / float(g["sigma2"]) < tol * 10
    assert _rel(r["trace"]["sigma2"], g["tr_sigma2"]) < tol * 10


@pytest.mark.parametrize("name", ["track_c1", "track_c1_b", "track_occl_head", "track_occl_mid", "track_all_visible"])
def test_oracle_reproduces_tracking_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    r = oracle.tracking_step(g["X"].astype(np.float64), g["Y_in"], 0.0, g["rest"], g["vis"], g["vis_ext"], oracle.TrackParams())
    assert r["state"] == int(g["state"]) and r["err"] == int(g["err"])
    assert list(r["iters"]) == list(g["iters"])
    assert r["priors"].shape == g["priors"].shape
    assert _rel(r["priors"], g["priors"]) < 1e-6
    assert _rel(r["Y"], g["Y"]) < 1e-6


@pytest.mark.parametrize("case", [
    dict(n_nodes=30, n_points=2000, occlusion=0.0, kw=dict(max_iter=20, tol=0.0)),
    dict(n_nodes=50, n_points=4000, occlusion=0.0, kw=dict()),
    dict(n_nodes=50, n_points=6000, occlusion=0.4, kw=dict(alpha=3.0, k_vis=50.0, visibility_threshold=0.008, max_iter=15, tol=0.0),
         priors=True, vis=True),
])
def test_oracle_matches_numpy_twin(case):
    f = synth.make_frame(7, n_nodes=case["n_nodes"], n_points=case["n_points"], occlusion=case["occlusion"])
    prm = oracle.CpdParams(**case["kw"])
    priors = None
    if case.get("priors"):
        sel = np.arange(0, case["n_nodes"], 5)
        priors = np.concatenate([sel[:, None].astype(float), f["Y"][sel] + 0.001], axis=1)
    vis = f["vis_ext"] if case.get("vis") else None
    a = oracle.cpd_lle(f["X"], f["Y"], 0.0, prm, priors=priors, vis=vis)
    b = nt.cpd_lle(f["X"], f["Y"], 0.0, prm.beta, prm.lambda_, prm.lle_weight, prm.mu, prm.max_iter, prm.tol,
                   priors=priors, alpha=prm.alpha, vis=vis, k_vis=prm.k_vis, tau=prm.visibility_threshold)
    assert a["iters"] == b["iters"] and a["converged"] == b["converged"] and a["kept"] == b["kept"]
    assert _rel(a["Y"], b["Y"]) < 1e-10
    assert _rel(a["W"], b["W"]) < 1e-8
    assert abs(a["sigma2"] - b["sigma2"]) / b["sigma2"] < 1e-9


def test_oracle_lle_twin_given_H():
    """include_lle path with H supplied: oracle and twin must agree (H itself is rounding noise)."""
    f = synth.make_frame(3, n_nodes=30, n_points=2000)
    H = oracle.lle_H(f["Y"])
    prm = oracle.CpdParams(beta=3.0, lambda_=1.0, include_lle=True, max_iter=10, tol=0.0)
    a = oracle.cpd_lle(f["X"], f["Y"], 0.0, prm, H=H)
    b = nt.cpd_lle(f["X"], f["Y"], 0.0, prm.beta, prm.lambda_, prm.lle_weight, prm.mu, prm.max_iter, prm.tol,
                   include_lle=True, H=H)
    assert _rel(a["Y"], b["Y"]) < 1e-8


def test_lle_rows_sum_to_one():
    f = synth.make_frame(0, n_nodes=30, n_points=100)
    H = oracle.lle_H(f["Y"])
    assert H.shape == (30, 30)
    assert np.allclose(H, H.T, atol=1e-6 * np.abs(H).max())
    # (I - L) 1 = 0  =>  H 1 = 0 up to the (large) rounding noise of the weights
    assert np.abs(H @ np.ones(30)).max() < 1e-6 * max(1.0, np.abs(H).max())


def test_max_iter_zero_returns_input_and_inits_sigma2():
    f = synth.make_frame(0, n_nodes=30, n_points=500)
    r = oracle.cpd_lle(f["X"], f["Y"], 0.0, oracle.CpdParams(max_iter=0))
    assert r["converged"] and r["iters"] == 0 and np.array_equal(r["Y"], f["Y"])
    d0 = np.linalg.norm(f["Y"][:, None] - f["X"][None], axis=2)
    Xk = f["X"][d0.min(0) < 0.1]
    expect = ((f["Y"][:, None] - Xk[None]) ** 2).sum() / (3 * 30 * len(Xk))
    assert abs(r["sigma2"] - expect) / expect < 1e-12


def test_prune_radius_is_strict():
    Y = np.stack([np.linspace(0, 0.3, 8), np.zeros(8), np.zeros(8)], axis=1)
    X = np.array([[0.0, 0.1, 0.0], [0.0, 0.0999, 0.0], [0.15, 0.05, 0.0], [5.0, 5.0, 5.0]])
    r = oracle.cpd_lle(X, Y, 0.0, oracle.CpdParams(max_iter=0))
    assert r["kept"] == 2          # 0.1 itself is rejected (trackdlo.cpp:190), the far point too


def test_end_node_substitution_quirk():
    """A point nearest to node 0 whose (substituted) neighbour 2 is closer than node 1 leaves row 1 at
    geodesic distance 0 => P=1 there (trackdlo.cpp:313-350).  Checked through P1 of the first iteration."""
    # hair-pin: node 2 folds back next to node 0, node 1 sticks out
    Y = np.array([[0.0, 0, 0], [0.03, 0.03, 0], [0.001, 0.004, 0], [0.0, 0.06, 0], [0.0, 0.09, 0], [0.0, 0.12, 0]])
    X = np.array([[0.0, -0.001, 0.0]])
    prm = oracle.CpdParams(max_iter=1, tol=0.0, mu=0.1)
    r = oracle.cpd_lle(X, Y, 1e-4, prm, trace=True)
    P1 = r["trace"]["P1"][0]
    # rows 0 (a), 2 (b) get exp(-d^2/2s2) ~ 1, row 1 gets exp(0) = 1 as well -> all three ~ equal
    assert P1[1] > 0.9 * P1[0] and P1[1] >= P1[2]


def test_all_underflow_column_contributes_nothing():
    f = synth.make_frame(0, n_nodes=30, n_points=300)
    far = np.array([[0.0, 0.0, 0.65 + 0.0995]])      # inside the 0.1 prune radius of some node? make sure it is kept
    X = np.concatenate([f["X"], far])
    a = oracle.cpd_lle(f["X"], f["Y"], 1e-7, oracle.CpdParams(max_iter=1, tol=0.0))
    b = oracle.cpd_lle(X, f["Y"], 1e-7, oracle.CpdParams(max_iter=1, tol=0.0))
    if b["kept"] == a["kept"] + 1:
        # the extra point's whole column underflows; c changes through N only
        assert np.isfinite(b["Y"]).all()


def test_traverse_straight_line_recovers_rest_spacing():
    n = 12
    guide = np.stack([np.linspace(0, 0.55, n), np.zeros(n), np.zeros(n)], axis=1)
    rest = np.linspace(0, 0.44, n)                   # rest spacing 0.04 (shorter than the guide's 0.05)
    vis = np.arange(n)
    pairs, err = oracle.traverse_euclidean(rest, guide, vis, 0)
    assert err == 0 and len(pairs) == n
    assert np.allclose(pairs[:, 0], np.arange(n))
    assert np.allclose(np.diff(pairs[:, 1]), 0.04, atol=1e-12)
    pairs1, err = oracle.traverse_euclidean(rest, guide, vis, 1)
    assert err == 0 and len(pairs1) == n
    assert np.allclose(pairs1[:, 0], np.arange(n)[::-1])
    assert np.allclose(np.diff(pairs1[:, 1]), -0.04, atol=1e-12)


def test_traverse_stops_at_visibility_gap():
    n = 12
    guide_full = np.stack([np.linspace(0, 0.55, n), np.zeros(n), np.zeros(n)], axis=1)
    vis = np.array([0, 1, 2, 3, 4, 8, 9, 10, 11])
    guide = guide_full[vis]
    rest = np.linspace(0, 0.55, n)
    head, _ = oracle.traverse_euclidean(rest, guide, vis, 0)
    tail, _ = oracle.traverse_euclidean(rest, guide, vis, 1)
    assert list(head[:, 0]) == [0, 1, 2, 3, 4]
    assert list(tail[:, 0]) == [11, 10, 9, 8]


def test_tracking_states():
    tp = oracle.TrackParams(max_iter=5)
    f = synth.make_frame(1, n_nodes=50, n_points=6000, occlusion=0.4)
    r = oracle.tracking_step(f["X"], f["Y"], 0.0, f["rest"], f["vis"], f["vis_ext"], tp)
    assert r["state"] == 1 and r["err"] == 0        # both ends visible, middle occluded
    n = 50
    for vis, state in ((np.arange(0, 30), 2), (np.arange(20, 50), 3), (np.arange(10, 40), 4)):
        r = oracle.tracking_step(f["X"], f["Y"], 0.0, f["rest"], vis.astype(np.int32), vis.astype(np.int32), tp)
        assert r["state"] == state
        assert np.isfinite(r["Y"]).all()


def test_visibility_front_end_oracle_matches_generator():
    """oracle_visibility (trackdlo_node.cpp:254-277, 346-360 restated in C++) against the independent NumPy
    statement in trackdlo_b200/synth.py, incl. occlusion gaps wider / narrower than d_vis and an empty cloud."""
    import oracle
    from trackdlo_b200 import synth
    for occ in (0.0, 0.05, 0.3, 0.55):
        f = synth.make_frame(7, n_nodes=50, n_points=4000, occlusion=occ)
        r = oracle.visibility(f["X"], f["Y"], f["rest"], 0.008, 0.06)
        assert np.array_equal(r["vis"], f["vis"]) and np.array_equal(r["vis_ext"], f["vis_ext"]), occ
        d = np.linalg.norm(f["Y"][:, None, :] - f["X"][None, :, :], axis=2).min(axis=1)
        assert np.allclose(r["dmin"], d, rtol=1e-14, atol=0)
    f = synth.make_frame(7, n_nodes=20, n_points=100)
    r = oracle.visibility(f["X"][:0], f["Y"], f["rest"])
    assert len(r["vis"]) == 0 and len(r["vis_ext"]) == 0 and np.all(r["dmin"] == 100000.0)


def test_evaluator_error_metric_oracle_known_answers():
    """oracle_tracking_error (evaluator.cpp:233-283, 333-341) on hand-computable cases."""
    import oracle
    line = np.array([[0.0, 0, 0], [1.0, 0, 0], [2.0, 0, 0]])
    assert oracle.tracking_error(line, line) == 0.0
    up = line + np.array([0.0, 0.5, 0.0])
    assert abs(oracle.tracking_error(up, line) - 0.5) < 1e-15            # parallel polylines 0.5 apart
    # nodes beyond the ends: distances to the end points, not to the infinite lines
    p = np.array([[3.0, 4.0, 0.0], [4.0, 4.0, 0.0]])
    e1 = (np.hypot(1.0, 4.0) + np.hypot(2.0, 4.0)) / 2                     # nodes of p -> end point (2,0,0)
    e2 = (5.0 + np.hypot(2.0, 4.0) + np.hypot(1.0, 4.0)) / 3               # nodes of line -> end point (3,4,0)
    assert abs(oracle.tracking_error(p, line) - (e1 + e2) / 2) < 1e-14
