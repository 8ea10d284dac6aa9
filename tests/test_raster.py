"""Self-occlusion test of the visibility block (raster part; trackdlo_node.cpp:280-343).
CPU: oracle/raster.py against OpenCV itself (cv2.line) and against the cv2-minted goldens.  GPU: tdlo_visibility_batched with a
projection matrix against the goldens and against the oracle on random self-crossing chains."""
import glob
import os

import numpy as np
import pytest

from oracle import raster
from trackdlo_b200 import synth

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "selfocc_*.npz")))

try:
    import cv2
except Exception:          # the GPU box may not have it: the goldens carry its outputs
    cv2 = None


def _random_segment(rng, mode):
    W, H = int(rng.integers(40, 260)), int(rng.integers(40, 200))
    th = int(rng.choice([40, 40, 40, 2, 3, 4, 5, 7, 12, 41, 39]))
    lo = th // 2 + 3 if mode == 0 else (-60 if mode == 1 else -300)
    if W - lo <= lo or H - lo <= lo:
        lo = 0
    a = (int(rng.integers(lo, W - lo)), int(rng.integers(lo, H - lo)))
    b = (int(rng.integers(lo, W - lo)), int(rng.integers(lo, H - lo)))
    return W, H, th, a, b


@pytest.mark.skipif(cv2 is None, reason="opencv-python not installed")
def test_thick_line_equals_cv2_line():
    rng = np.random.default_rng(21)
    for t in range(1500):
        W, H, th, a, b = _random_segment(rng, t % 3)
        if t % 17 == 0:
            b = (a[0] + int(rng.integers(-2, 3)), a[1] + int(rng.integers(-2, 3)))        # (near-)degenerate segments
        ref = np.zeros((H, W), np.uint8); cv2.line(ref, a, b, 255, th)
        img = np.zeros((H, W), np.uint8); raster.thick_line(img, a, b, th)
        assert (ref == img).all(), (W, H, a, b, th)


@pytest.mark.skipif(cv2 is None, reason="opencv-python not installed")
def test_pixel_predicate_equals_cv2_raster():
    rng = np.random.default_rng(22)
    n = 0
    for t in range(400):
        W, H, th, a, b = _random_segment(rng, t % 3)
        ref = np.zeros((H, W), np.uint8); cv2.line(ref, a, b, 255, th)
        k = np.ones((3, 3), np.uint8)
        ys, xs = np.nonzero(cv2.dilate(ref, k) != cv2.erode(ref, k))                      # the boundary band, both sides
        sel = rng.permutation(len(ys))[:30]
        pts = [(int(xs[i]), int(ys[i])) for i in sel] + [(int(rng.integers(0, W)), int(rng.integers(0, H))) for _ in range(8)]
        for x, y in pts:
            assert raster.covers(x, y, W, H, a, b, th) == bool(ref[y, x]), (W, H, a, b, th, x, y)
            n += 1
    assert n > 8000


def test_circle_known_answer():
    # filled cv::Circle (midpoint algorithm), radius 20 = dlo_pixel_width 40 and radius 3: half-width of the span at row offset k.
    # (The full width is only reached on the centre row and the top / bottom rows are single pixels: OpenCV's circle.)
    assert raster.circle_halfwidths(20) == [20, 19, 19, 19, 19, 19, 19, 18, 18, 17, 17, 16, 16, 15, 14, 13, 12, 10, 8, 6, 0]
    assert raster.circle_halfwidths(3) == [3, 2, 2, 0]


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_matches_cv2_minted_golden(path):
    g = np.load(path)
    free = raster.self_occlusion(g["Y"], g["proj"], int(g["rows"]), int(g["cols"]), int(g["pixel_width"]))
    assert (free.astype(np.int32) == g["cv2_not_self_occluded"]).all()
    assert (np.array(raster.project_pixels(g["Y"], g["proj"])) == g["pixels"]).all()
    # the literal raster loop with the oracle's own rasteriser gives the same flags
    free_r = raster.self_occlusion(g["Y"], g["proj"], int(g["rows"]), int(g["cols"]), int(g["pixel_width"]), use_raster=True)
    assert (free_r == free).all()


def test_goldens_present():
    assert len(GOLD) >= 5


# ------------------------------------------------------------------------------------------------ GPU
def _crossing_chain(rng, n, rows, cols):
    """Random self-crossing chain in camera coordinates whose projection wanders over (and a little outside) the image."""
    P = synth.camera_matrix(rows, cols, f=915.0 * cols / 1280.0)
    t = np.linspace(0, 1, n)
    k1, k2 = rng.uniform(0.6, 2.2), rng.uniform(0.8, 3.0)
    amp = rng.uniform(0.35, 0.75)
    z = rng.uniform(0.45, 0.8) + 0.08 * np.cos(2 * np.pi * k1 * t + rng.uniform(0, 6)) + 0.03 * t
    u = cols / 2 + amp * cols * np.sin(2 * np.pi * k1 * t + rng.uniform(0, 6)) * (0.5 + 0.5 * t)
    v = rows / 2 + amp * rows * np.sin(2 * np.pi * k2 * t + rng.uniform(0, 6))
    x = (u - P[0, 2]) * z / P[0, 0]; y = (v - P[1, 2]) * z / P[1, 1]
    return np.stack([x, y, z], 1), P


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_gpu_self_occlusion_golden(path):
    from trackdlo_b200 import api
    g = np.load(path)
    Y, X = g["Y"], g["X"]; N = len(Y)
    ctx = api.Context(max_frames=1, max_nodes=N, max_points_total=len(X))
    r = ctx.visibility_batched(X, np.array([0, len(X)], np.int64), Y[None], g["node_coord"][None], float(g["visibility_threshold"]), float(g["d_vis"]),
                               proj=g["proj"][None], rows=int(g["rows"]), cols=int(g["cols"]), pixel_width=int(g["pixel_width"]))
    assert (r["not_self_occluded"][0] == g["cv2_not_self_occluded"]).all()
    assert list(r["visible"]) == list(g["visible"]) and list(r["visible_ext"]) == list(g["visible_ext"])
    # without the projection matrix the test is off: every node with a point nearby is visible
    r0 = ctx.visibility_batched(X, np.array([0, len(X)], np.int64), Y[None], g["node_coord"][None], float(g["visibility_threshold"]), float(g["d_vis"]))
    assert len(r0["visible"]) >= len(r["visible"]) and (r0["not_self_occluded"] == 1).all()
    ctx.close()


@pytest.mark.gpu
def test_gpu_self_occlusion_random_chains_batched():
    from trackdlo_b200 import api
    rng = np.random.default_rng(5)
    occluded = 0
    for rnd in range(6):
        F, N = 16, int(rng.integers(8, 70))
        rows, cols = int(rng.choice([120, 240, 480, 720])), int(rng.choice([160, 320, 640, 1280]))
        width = int(rng.choice([40, 40, 25, 7, 2, 41, 64]))
        Ys, Ps = zip(*[_crossing_chain(rng, N, rows, cols) for _ in range(F)])
        Y = np.stack(Ys); P = np.stack(Ps)
        X = np.concatenate([y + rng.normal(0, 0.001, y.shape) for y in Ys]); xo = np.arange(F + 1, dtype=np.int64) * N
        nc = np.stack([np.concatenate([[0], np.cumsum(np.linalg.norm(np.diff(y, axis=0), axis=1))]) for y in Ys])
        ctx = api.Context(max_frames=F, max_nodes=N, max_points_total=len(X))
        r = ctx.visibility_batched(X, xo, Y, nc, 0.008, 0.06, proj=P, rows=rows, cols=cols, pixel_width=width)
        for f in range(F):
            free = raster.self_occlusion(Ys[f], Ps[f], rows, cols, width)
            assert (r["not_self_occluded"][f] == free.astype(np.int32)).all(), (rnd, f, N, rows, cols, width)
            vis = [i for i in range(N) if free[i]]                       # every node has a point within the threshold
            assert list(r["visible"][r["visible_offsets"][f]:r["visible_offsets"][f + 1]]) == vis
            occluded += int((~free).sum())
        ctx.close()
    assert occluded > 50          # the chains do cross


@pytest.mark.gpu
def test_gpu_self_occlusion_argument_checks():
    from trackdlo_b200 import api
    g = np.load(GOLD[0]); Y, X = g["Y"], g["X"]
    ctx = api.Context(max_frames=1, max_nodes=len(Y), max_points_total=len(X))
    for kw in (dict(rows=0, cols=100, pixel_width=40), dict(rows=100, cols=100, pixel_width=1), dict(rows=100, cols=100, pixel_width=2000)):
        with pytest.raises(api.TdloError):
            ctx.visibility_batched(X, np.array([0, len(X)], np.int64), Y[None], g["node_coord"][None], proj=g["proj"][None], **kw)
    ctx.close()


@pytest.mark.gpu
def test_gpu_sequence_mode_with_self_occlusion():
    """Sequence mode with a camera: every step's visibility lists carry the self-occlusion test -- same trajectory as chaining
    tdlo_visibility_batched(proj) and tdlo_tracking_step_batched by hand, and different from the run without the camera."""
    from trackdlo_b200 import api
    g = np.load([p for p in GOLD if p.endswith("selfocc_coil.npz")][0])
    Y0, X, P = g["Y"], g["X"], g["proj"]; N = len(Y0)
    rows, cols, width = int(g["rows"]), int(g["cols"]), int(g["pixel_width"])
    T = 3
    rng = np.random.default_rng(9)
    clouds = [X + rng.normal(0, 0.0005, X.shape) + np.array([0.002 * t, 0.0, 0.0]) for t in range(T)]
    xo = np.zeros(T + 1, np.int64); xo[1:] = np.cumsum([len(c) for c in clouds])
    rest = g["node_coord"]
    tp = api.TrackParams(max_iter=8)
    ctx = api.Context(max_frames=1, max_nodes=N, max_points_total=len(X))
    r = ctx.track_sequences(np.concatenate(clouds), xo, Y0[None], np.zeros(1), rest[None], tp, T, d_vis=0.06, proj=P[None], rows=rows, cols=cols, pixel_width=width)
    r0 = ctx.track_sequences(np.concatenate(clouds), xo, Y0[None], np.zeros(1), rest[None], tp, T, d_vis=0.06)
    one = lambda n: np.array([0, n], np.int64)
    Y, s2 = Y0.copy(), np.zeros(1)
    hidden = 0
    for t in range(T):
        v = ctx.visibility_batched(clouds[t], one(len(clouds[t])), Y[None], rest[None], tp.visibility_threshold, 0.06, proj=P[None], rows=rows, cols=cols, pixel_width=width)
        hidden += int((v["not_self_occluded"] == 0).sum())
        o = ctx.tracking_step_batched(clouds[t], one(len(clouds[t])), Y[None], s2, rest[None], v["visible"], v["visible_offsets"], v["visible_ext"], v["visible_ext_offsets"], tp)
        Y, s2 = o["Y"][0], o["sigma2"]
        assert list(r["iters"][t, 0]) == list(o["iters"][0])
        assert np.array_equal(r["Y_traj"][t, 0], Y)                  # same kernels, same inputs: bit-identical
    assert hidden > 0
    assert not np.array_equal(r["Y"], r0["Y"])                       # the occluded nodes changed the registration
    ctx.close()
