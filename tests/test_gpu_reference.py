"""GPU parity, part 2: the CUDA path (through the C ABI) against

  * the REFERENCE's own outputs: the `ref_*` fields of tests/golden/*.npz (minted here by scripts/make_golden.py from
    oracle/_ref/libtrackdlo_ref.so = the unmodified trackdlo.cpp + utils.cpp) and, when the prebuilt library travelled to
    the box, the reference run live on the same inputs;
  * the oracle at the FULL sizes bench.py times (BASELINE configs[1], [3], [4]: C2 / C4 / C5 as tracking_step);
  * a bounded, fixed-seed slice of the randomised differential scripts (scripts/fuzz_*.py), so that the driver re-runs it.

Tolerances as in test_gpu_parity.py: gate 1e-5 relative (north_star); held: 1e-6 on tracking_step outputs, integer
outputs (iteration counts, states, prior indices, status) exact."""
import glob
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import oracle
from oracle import ref
from trackdlo_b200 import api, synth

pytestmark = pytest.mark.gpu

GATE = 1e-5
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TRACK_GOLDENS = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(ROOT, "tests", "golden", "track_*.npz")))


def rel(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)) if a.size else 0.0


def one(n):
    return np.array([0, n], np.int64)


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(max_frames=64, max_nodes=64, max_points_total=64 * 20000 + 1000)
    yield c
    c.close()


def _track(c, f, tp, s2=0.0):
    return c.tracking_step_batched(f["X"], one(len(f["X"])), f["Y"][None], np.array([s2]), f["rest"][None], f["vis"], one(len(f["vis"])),
                                   f["vis_ext"], one(len(f["vis_ext"])), tp)


# ------------------------------------------------------------------------------------------------
# reference outputs carried by the goldens (all five tracking_step states, incl. 2 and 4 -> alignment 2)
# ------------------------------------------------------------------------------------------------
def test_goldens_cover_every_tracking_state(golden_dir):
    states = {int(np.load(os.path.join(golden_dir, n + ".npz"))["ref_state"]) for n in TRACK_GOLDENS}
    assert states == {0, 1, 2, 3, 4}


@pytest.mark.parametrize("name", TRACK_GOLDENS)
@pytest.mark.parametrize("chunk", [0, 256])
def test_tracking_step_against_reference_outputs(ctx, golden_dir, name, chunk):
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    f = dict(X=g["X"].astype(np.float64), Y=g["Y_in"], rest=g["rest"], vis=g["vis"].astype(np.int32), vis_ext=g["vis_ext"].astype(np.int32))
    ctx.set_option("chunk_points", chunk)
    try:
        r = _track(ctx, f, api.TrackParams())
    finally:
        ctx.set_option("chunk_points", 0)
    npri = len(g["ref_priors"])
    assert r["status"][0] == 0
    assert r["state"][0] == int(g["ref_state"]) == int(g["state"])
    assert list(r["iters"][0]) == list(g["ref_iters"])
    assert r["n_priors"][0] == npri
    assert np.array_equal(r["priors"][0, :npri, 0], g["ref_priors"][:, 0])            # node indices of traverse_euclidean: exact
    assert rel(r["priors"][0, :npri, 1:], g["ref_priors"][:, 1:]) < 1e-6
    assert rel(r["guide"][0, :len(f["vis_ext"])], g["ref_guide"]) < 1e-6
    assert rel(r["Y"][0], g["ref_Y"]) < 1e-6 < GATE
    assert abs(r["sigma2"][0] - float(g["ref_sigma2"])) / float(g["ref_sigma2"]) < 1e-5


@pytest.mark.parametrize("name", ["c1_fixed20", "c1_converge", "c1_lle_preproc", "occl_vis_priors", "n64_sigma_given"])
def test_cpd_against_reference_outputs(ctx, golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"cpd_{name}.npz"))
    a = g["params"]
    pg = api.CpdParams(beta=a[0], lambda_=a[1], lle_weight=a[2], mu=a[3], tol=a[4], alpha=a[5], k_vis=a[6], visibility_threshold=a[7],
                       max_iter=int(a[8]), include_lle=bool(a[9]))
    X = g["X"].astype(np.float64); Nn = g["Y_in"].shape[0]
    pri = npr = nv = None
    if len(g["priors"]):
        pri = np.zeros((1, Nn, 4)); pri[0, :len(g["priors"])] = g["priors"]; npr = np.array([len(g["priors"])], np.int32)
    if int(g["n_visible"]) >= 0:
        nv = np.array([int(g["n_visible"])], np.int32)
    r = ctx.cpd_lle_batched(X, one(len(X)), g["Y_in"][None], np.array([float(g["sigma2_in"])]), pg, priors=pri, n_priors=npr, n_visible=nv)
    assert r["iters"][0] == int(g["ref_iters"])
    assert bool(r["status"][0] & api.ST_NOT_CONVERGED) == (not bool(g["ref_converged"]))
    tol = 1e-7 if not pg.include_lle else 1e-6
    assert rel(r["Y"][0], g["ref_Y"]) < tol < GATE
    assert abs(r["sigma2"][0] - float(g["ref_sigma2"])) / float(g["ref_sigma2"]) < tol * 10


# ------------------------------------------------------------------------------------------------
# the reference run live on the box (the prebuilt oracle/_ref travels with the snapshot)
# ------------------------------------------------------------------------------------------------
needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libtrackdlo_ref.so did not travel")

STATE_WINDOWS = {0: None, 1: [(0.35, 0.65)], 2: [(0.7, 1.0)], 3: [(0.0, 0.3)], 4: [(0.0, 0.2), (0.8, 1.0)]}


@needs_ref
@pytest.mark.parametrize("state", [0, 1, 2, 3, 4])
def test_tracking_step_states_against_live_reference(ctx, state):
    """Every branch of trackdlo.cpp:929-995; for state 4 three different occlusion layouts so that the alignment node
    and both walking directions of traverse_euclidean alignment 2 (:749-895) vary."""
    layouts = [STATE_WINDOWS[state]] if state != 4 else [[(0.0, 0.2), (0.8, 1.0)], [(0.0, 0.15), (0.45, 0.6), (0.85, 1.0)], [(0.0, 0.3), (0.5, 0.55), (0.9, 1.0)]]
    for li, win in enumerate(layouts):
        for idx in range(3):
            kw = dict(n_nodes=40, n_points=3000, occl_windows=win)
            if state == 0:
                kw["tau_vis"] = 0.02
            f = synth.make_frame(60 + 10 * li + idx, **kw)
            if state in (1, 2) and f["vis_ext"][0] != 0:            # the end node itself may sit just outside tau_vis
                f["vis"] = np.concatenate([[0], f["vis"]]).astype(np.int32); f["vis_ext"] = np.concatenate([[0], f["vis_ext"]]).astype(np.int32)
            if state in (1, 3) and f["vis_ext"][-1] != 39:
                f["vis"] = np.concatenate([f["vis"], [39]]).astype(np.int32); f["vis_ext"] = np.concatenate([f["vis_ext"], [39]]).astype(np.int32)
            o = oracle.tracking_step(f["X"], f["Y"], 0.0, f["rest"], f["vis"], f["vis_ext"], oracle.TrackParams())
            if o["err"] != 0:
                continue                        # the reference itself reads out of range on this input
            rr = ref.tracking_step(f["X"], f["Y"], 0.0, f["rest"], f["vis"], f["vis_ext"], oracle.TrackParams())
            r = _track(ctx, f, api.TrackParams())
            assert r["status"][0] & ~(api.ST_NOT_CONVERGED | api.ST_PRE_NOT_CONVERGED) == 0 and r["state"][0] == rr["state"] == state
            assert list(r["iters"][0]) == list(rr["iters"])
            npri = len(rr["priors"])
            assert r["n_priors"][0] == npri and np.array_equal(r["priors"][0, :npri, 0], rr["priors"][:, 0])
            assert rel(r["priors"][0, :npri, 1:], rr["priors"][:, 1:]) < 1e-6
            assert rel(r["Y"][0], rr["Y"]) < 1e-6 < GATE


@needs_ref
def test_random_cpd_lle_against_live_reference(ctx):
    rng = np.random.default_rng(77)
    for case in range(20):
        Nn = int(rng.integers(6, 65)); Mp = int(rng.integers(100, 4000))
        f = synth.make_frame(5000 + case, n_nodes=Nn, n_points=Mp, occlusion=float(rng.choice([0.0, 0.3])), occl_start=float(rng.uniform(0, 0.6)))
        kw = dict(max_iter=int(rng.integers(1, 20)), tol=float(rng.choice([0.0, 2e-4])), mu=float(rng.choice([0.05, 0.1])))
        pri = npr = nv = None; vis = None; priors = None
        if rng.integers(0, 2):
            k = int(rng.integers(1, Nn)); sel = np.sort(rng.choice(Nn, size=k, replace=False))
            priors = np.concatenate([sel[:, None].astype(float), f["Y"][sel] + rng.normal(0, 0.003, (k, 3))], axis=1)
            pri = np.zeros((1, Nn, 4)); pri[0, :k] = priors; npr = np.array([k], np.int32); kw["alpha"] = 3.0
        if rng.integers(0, 2) and 0 < len(f["vis_ext"]) < Nn:
            vis = np.arange(len(f["vis_ext"])); nv = np.array([len(vis)], np.int32); kw["k_vis"] = 50.0; kw["visibility_threshold"] = 0.008
        s2 = float(rng.choice([0.0, 1e-4]))
        rr = ref.cpd_lle(f["X"], f["Y"], s2, oracle.CpdParams(**kw), priors=priors, vis=vis)
        r = ctx.cpd_lle_batched(f["X"], one(len(f["X"])), f["Y"][None], np.array([s2]), api.CpdParams(**kw), priors=pri, n_priors=npr, n_visible=nv)
        assert r["iters"][0] == rr["iters"], (case, kw)
        assert bool(r["status"][0] & api.ST_NOT_CONVERGED) == (not rr["converged"])
        assert rel(r["Y"][0], rr["Y"]) < 1e-7, (case, kw)
        assert abs(r["sigma2"][0] - rr["sigma2"]) / rr["sigma2"] < 1e-6


# ------------------------------------------------------------------------------------------------
# what bench.py times, at full size
# ------------------------------------------------------------------------------------------------
def _oracle_track_many(frames, tp, s2=0.0):
    with ThreadPoolExecutor(max_workers=min(len(frames), os.cpu_count() or 1)) as ex:       # ctypes releases the GIL
        return list(ex.map(lambda f: oracle.tracking_step(f["X"], f["Y"], s2, f["rest"], f["vis"], f["vis_ext"], tp), frames))


def test_c2_full_batch_tracking_step(ctx):
    """BASELINE configs[1] exactly as bench.py runs it: 64 frames x (Nn=50, Mp=20000), full tracking_step, max_iter=50,
    tol=0 (100 EM iterations per frame).  Frames 0, 31 and 63 of the batch against the oracle."""
    wl = synth.make_batch(64, n_nodes=50, n_points=20000)
    tp = api.TrackParams(max_iter=50, tol=0.0)
    r = ctx.tracking_step_batched(wl["X"], wl["x_offsets"], wl["Y"], np.zeros(64), wl["rest"], wl["vis"], wl["vis_offsets"],
                                  wl["vis_ext"], wl["vis_ext_offsets"], tp)
    assert np.all(r["iters"] == 50)
    assert np.all(r["status"] & ~(api.ST_NOT_CONVERGED | api.ST_PRE_NOT_CONVERGED) == 0)
    pick = [0, 31, 63]
    outs = _oracle_track_many([wl["frames"][i] for i in pick], oracle.TrackParams(max_iter=50, tol=0.0))
    for i, o in zip(pick, outs):
        assert o["err"] == 0 and r["state"][i] == o["state"]
        npri = len(o["priors"])
        assert r["n_priors"][i] == npri and np.array_equal(r["priors"][i, :npri, 0], o["priors"][:, 0])
        assert rel(r["Y"][i], o["Y"]) < 1e-6 < GATE, i
        assert abs(r["sigma2"][i] - o["sigma2"]) / o["sigma2"] < 1e-5


def test_c4_shape_default_tolerance_iteration_counts(ctx):
    """BASELINE configs[3] shape: frames of (Nn=50, Mp=20000) at the DEFAULT tol/max_iter (data-dependent iteration
    counts, trackdlo.cpp:424-428).  64 frames; the iteration counts of both registrations must equal the oracle's in
    every frame, Y within the gate."""
    F = 64
    wl = synth.make_batch(F, first_frame=1000, n_nodes=50, n_points=20000)
    tp = api.TrackParams()
    r = ctx.tracking_step_batched(wl["X"], wl["x_offsets"], wl["Y"], np.zeros(F), wl["rest"], wl["vis"], wl["vis_offsets"],
                                  wl["vis_ext"], wl["vis_ext_offsets"], tp)
    outs = _oracle_track_many(wl["frames"], oracle.TrackParams())
    for i, o in enumerate(outs):
        assert o["err"] == 0
        assert list(r["iters"][i]) == list(o["iters"]), i
        assert r["state"][i] == o["state"], i
        assert bool(r["status"][i] & api.ST_NOT_CONVERGED) == (not o["converged"][1]), i
        assert bool(r["status"][i] & api.ST_PRE_NOT_CONVERGED) == (not o["converged"][0]), i
        assert r["status"][i] & ~(api.ST_NOT_CONVERGED | api.ST_PRE_NOT_CONVERGED) == 0, i
        assert rel(r["Y"][i], o["Y"]) < 1e-6 < GATE, i


def test_c5_full_size_tracking_step():
    """BASELINE configs[4] shape (Nn=200, Mp=100000) as a full tracking_step: the pre-processing registration with the LLE
    regulariser (non-symmetric system at Nn=200) + traversal + main registration (SPD path), 3 iterations each."""
    Nn, Mp = 200, 100000
    f = synth.make_frame(3, n_nodes=Nn, n_points=Mp)
    c = api.Context(max_frames=1, max_nodes=Nn, max_points_total=Mp)
    try:
        tp = api.TrackParams(max_iter=3, tol=0.0)
        r = _track(c, f, tp)
        o = oracle.tracking_step(f["X"], f["Y"], 0.0, f["rest"], f["vis"], f["vis_ext"], oracle.TrackParams(max_iter=3, tol=0.0))
        assert o["err"] == 0 and r["state"][0] == o["state"] and list(r["iters"][0]) == [3, 3]
        npri = len(o["priors"])
        assert r["n_priors"][0] == npri and np.array_equal(r["priors"][0, :npri, 0], o["priors"][:, 0])
        assert rel(r["guide"][0, :len(f["vis_ext"])], o["guide"]) < 1e-6
        assert rel(r["Y"][0], o["Y"]) < 1e-6 < GATE
    finally:
        c.close()


# ------------------------------------------------------------------------------------------------
# bounded slice of scripts/fuzz_parity.py / fuzz_batches.py (fixed seeds)
# ------------------------------------------------------------------------------------------------
def test_fuzz_slice_single_frames(ctx):
    rng = np.random.default_rng(1)
    n = 0
    for case in range(70):
        Nn = int(rng.integers(4, 65)); Mp = int(rng.integers(50, 9000)); occ = float(rng.choice([0.0, 0.0, 0.2, 0.45]))
        start = float(rng.choice([0.3, 0.0, 0.6, 0.75]))
        f = synth.make_frame(int(rng.integers(0, 10000)), n_nodes=Nn, n_points=Mp, occlusion=occ, occl_start=start)
        ctx.set_option("chunk_points", int(rng.choice([0, 256, 512, 1024, 4096])))
        ctx.set_option("truncation", float(rng.choice([100.0, 745.2])))
        ctx.set_option("truncation_rel", float(rng.choice([45.0, 745.2])))
        ctx.set_option("threads", int(rng.choice([256, 256])))
        mi = int(rng.integers(1, 25)); tol = float(rng.choice([0.0, 2e-4]))
        try:
            if len(f["vis_ext"]) < 4 or len(f["X"]) == 0:
                continue      # < 4 guide nodes: out-of-range reads in the reference (trackdlo.cpp:92-117, 313-321)
            if rng.random() < 0.4:
                kw = dict(max_iter=mi, tol=tol, include_lle=bool(rng.random() < 0.3))
                if kw["include_lle"]:
                    kw.update(beta=3.0, lambda_=1.0)
                o = oracle.cpd_lle(f["X"], f["Y"], 0.0, oracle.CpdParams(**kw))
                r = ctx.cpd_lle_batched(f["X"], one(len(f["X"])), f["Y"][None], np.zeros(1), api.CpdParams(**kw))
                assert r["iters"][0] == o["iters"], (case, kw)
            else:
                o = oracle.tracking_step(f["X"], f["Y"], 0.0, f["rest"], f["vis"], f["vis_ext"], oracle.TrackParams(max_iter=mi, tol=tol))
                if o["err"] != 0:
                    continue
                r = _track(ctx, f, api.TrackParams(max_iter=mi, tol=tol))
                assert list(r["iters"][0]) == list(o["iters"]) and r["state"][0] == o["state"], (case, Nn, Mp, occ, start)
            assert rel(r["Y"][0], o["Y"]) < 1e-6, (case, Nn, Mp, occ, start)
            n += 1
        finally:
            ctx.set_option("chunk_points", 0); ctx.set_option("truncation", 100.0); ctx.set_option("truncation_rel", 45.0); ctx.set_option("threads", 256)
    assert n >= 50


def test_fuzz_slice_ragged_batches(ctx):
    rng = np.random.default_rng(3)
    checked = 0
    for case in range(12):
        F = int(rng.integers(1, 7)); Nn = int(rng.integers(8, 65))
        frames = []
        while len(frames) < F:
            f = synth.make_frame(int(rng.integers(0, 100000)), n_nodes=Nn, n_points=int(rng.integers(100, 7000)),
                                 occlusion=float(rng.choice([0.0, 0.15, 0.4])), occl_start=float(rng.choice([0.3, 0.0, 0.7])))
            if len(f["vis_ext"]) >= 4:
                frames.append(f)
        ctx.set_option("chunk_points", int(rng.choice([0, 256, 1024, 2048]))); ctx.set_option("threads", int(rng.choice([256, 256])))
        try:
            s2 = np.where(rng.random(F) < 0.3, 10.0 ** rng.uniform(-6, -3, F), 0.0)
            mi = int(rng.integers(1, 20)); tol = float(rng.choice([0.0, 2e-4]))
            xo = np.zeros(F + 1, np.int64); xo[1:] = np.cumsum([len(f["X"]) for f in frames])
            vo = np.zeros(F + 1, np.int64); vo[1:] = np.cumsum([len(f["vis"]) for f in frames])
            eo = np.zeros(F + 1, np.int64); eo[1:] = np.cumsum([len(f["vis_ext"]) for f in frames])
            r = ctx.tracking_step_batched(np.concatenate([f["X"] for f in frames]), xo, np.stack([f["Y"] for f in frames]), s2,
                                          np.stack([f["rest"] for f in frames]), np.concatenate([f["vis"] for f in frames]), vo,
                                          np.concatenate([f["vis_ext"] for f in frames]), eo, api.TrackParams(max_iter=mi, tol=tol))
            for i, f in enumerate(frames):
                o = oracle.tracking_step(f["X"], f["Y"], float(s2[i]), f["rest"], f["vis"], f["vis_ext"], oracle.TrackParams(max_iter=mi, tol=tol))
                if o["err"] != 0:
                    continue
                assert list(r["iters"][i]) == list(o["iters"]) and r["state"][i] == o["state"] and r["n_priors"][i] == len(o["priors"]), (case, i)
                assert rel(r["Y"][i], o["Y"]) < 1e-6 and abs(r["sigma2"][i] - o["sigma2"]) / o["sigma2"] < 1e-5, (case, i)
                checked += 1
        finally:
            ctx.set_option("chunk_points", 0); ctx.set_option("threads", 256)
    assert checked >= 25
