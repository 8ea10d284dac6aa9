"""Seeded synthetic DLO frames: the workload generator for tests and bench.py.

Mirrors the data shapes the reference's front-end hands to trackdlo::tracking_step
(trackdlo_node.cpp:242 float32-origin points; :254-277, :346-360 visibility lists).
"""
import numpy as np

L_CURVE = 0.8
SEED0 = 20231010


def curve(t):
    t = np.asarray(t, dtype=np.float64)
    return np.stack([L_CURVE * (t - 0.5), 0.08 * np.sin(2 * np.pi * t), 0.65 + 0.03 * np.cos(3 * t)], axis=-1)


def observed_curve(t, frame_idx):
    t = np.asarray(t, dtype=np.float64)
    ph = 0.1 * frame_idx
    off = 0.01 * np.stack([np.sin(3 * t + ph), np.cos(2 * t + ph), np.zeros_like(t)], axis=-1)
    return curve(t) + off


def rest_arclengths(Y):
    seg = np.linalg.norm(np.diff(Y, axis=0), axis=1)
    return np.concatenate([[0.0], np.cumsum(seg)])


def visibility(Y, X, rest, tau_vis=0.008, d_vis=0.06):
    """visible_nodes / visible_nodes_extended as built by trackdlo_node.cpp:254-277,346-360
    (self-occlusion raster excluded)."""
    Nn = Y.shape[0]
    dmin = np.full(Nn, np.inf)
    for lo in range(0, X.shape[0], 65536):
        d = np.linalg.norm(Y[:, None, :] - X[None, lo:lo + 65536, :], axis=2)
        dmin = np.minimum(dmin, d.min(axis=1))
    vis = [int(m) for m in range(Nn) if dmin[m] <= tau_vis]
    ext = []
    for i in range(len(vis) - 1):
        ext.append(vis[i])
        if abs(rest[vis[i + 1]] - rest[vis[i]]) <= d_vis:
            ext.extend(range(vis[i] + 1, vis[i + 1]))
    if vis:
        ext.append(vis[-1])
    return np.asarray(vis, np.int32), np.asarray(ext, np.int32)


def make_frame(frame_idx, n_nodes=50, n_points=20000, occlusion=0.0, tau_vis=0.008, d_vis=0.06,
               outlier_frac=0.01, noise=0.002, occl_start=0.3, occl_windows=None):
    """One independent frame: dict(X [Mp,3] f64, Y [Nn,3], rest [Nn], vis, vis_ext).

    Occlusion deletes the points whose curve parameter lies in [occl_start, occl_start + occlusion] (SURVEY §8d:
    occl_start = 0.3 -> mid-section / head states of trackdlo.cpp:929-995); `occl_windows` = [(t0, t1), ...] deletes
    several windows instead, e.g. [(0.75, 1.0)] -> "Tail occluded", [(0, 0.2), (0.8, 1)] -> "Both ends occluded"."""
    rng = np.random.default_rng(SEED0 + frame_idx)
    Y = curve(np.linspace(0.0, 1.0, n_nodes))
    t = rng.random(n_points)
    X = observed_curve(t, frame_idx) + rng.normal(0.0, noise, size=(n_points, 3))
    n_out = int(round(outlier_frac * n_points))
    if n_out:
        idx = rng.choice(n_points, size=n_out, replace=False)
        X[idx] = np.array([0.0, 0.0, 0.65]) + rng.uniform(-0.3, 0.3, size=(n_out, 3))
    if occl_windows is None and occlusion > 0:
        occl_windows = [(occl_start, occl_start + occlusion)]
    if occl_windows:
        keep = np.ones(n_points, bool)
        for (t0, t1) in occl_windows:
            keep &= ~((t >= t0) & (t <= t1))
        X = X[keep]
    X = X.astype(np.float32).astype(np.float64)      # trackdlo_node.cpp:242
    rest = rest_arclengths(Y)
    vis, ext = visibility(Y, X, rest, tau_vis, d_vis)
    return dict(X=np.ascontiguousarray(X), Y=Y, rest=rest, vis=vis, vis_ext=ext)


def make_batch(n_frames, first_frame=0, **kw):
    """Ragged batch in the C-ABI layout: X concatenated [sum Mp, 3], x_offsets [F+1], Y [F,Nn,3], ..."""
    frames = [make_frame(first_frame + f, **kw) for f in range(n_frames)]
    offs = np.zeros(n_frames + 1, np.int64)
    offs[1:] = np.cumsum([f["X"].shape[0] for f in frames])
    voff = np.zeros(n_frames + 1, np.int64)
    voff[1:] = np.cumsum([len(f["vis"]) for f in frames])
    eoff = np.zeros(n_frames + 1, np.int64)
    eoff[1:] = np.cumsum([len(f["vis_ext"]) for f in frames])
    return dict(
        frames=frames,
        X=np.ascontiguousarray(np.concatenate([f["X"] for f in frames], axis=0)),
        x_offsets=offs,
        Y=np.ascontiguousarray(np.stack([f["Y"] for f in frames])),
        rest=np.ascontiguousarray(np.stack([f["rest"] for f in frames])),
        vis=np.concatenate([f["vis"] for f in frames]).astype(np.int32), vis_offsets=voff,
        vis_ext=np.concatenate([f["vis_ext"] for f in frames]).astype(np.int32), vis_ext_offsets=eoff,
    )


# ---------------------------------------------------------------------------------------------
# camera frames for the perception front-end (trackdlo_node.cpp:159-242): a blue DLO on a grey table
# ---------------------------------------------------------------------------------------------
def camera_matrix(rows=720, cols=1280, f=915.0):
    """3x4 projection matrix of a RealSense-like colour camera (launch/realsense_node.launch: 1280x720)."""
    P = np.zeros((3, 4)); P[0, 0] = f; P[1, 1] = f; P[0, 2] = cols / 2.0 - 0.5; P[1, 2] = rows / 2.0 - 0.5; P[2, 2] = 1.0
    return P


def render_frame(frame_idx, rows=720, cols=1280, width_px=7, occl_windows=None, occlusion_box=None, dlo_bgr=(200, 70, 20), depth_holes=0.01):
    """Synthetic colour + depth image of the observed curve of frame `frame_idx`: dict(bgr [H,W,3] uint8, depth [H,W] uint16 mm,
    proj [3,4], occlusion_bgr [H,W,3] or None).  The DLO is drawn as discs of `width_px` pixels around the projections of a dense
    sampling of observed_curve; `depth_holes` of its pixels get depth 0 (invalid RealSense returns, kept by the reference);
    occlusion_box = (i0, i1, j0, j1) blacks a rectangle of the occlusion image (utils/simulate_occlusion.py)."""
    rng = np.random.default_rng(SEED0 + 7919 * frame_idx)
    P = camera_matrix(rows, cols, f=915.0 * cols / 1280.0)
    bgr = np.empty((rows, cols, 3), np.uint8)
    bgr[:] = (118, 120, 123)                                     # table: low saturation
    bgr = np.clip(bgr.astype(np.int16) + rng.integers(-6, 7, bgr.shape), 0, 255).astype(np.uint8)
    depth = np.full((rows, cols), 1100, np.uint16)
    t = np.linspace(0.0, 1.0, 6000)
    if occl_windows:
        keep = np.ones(len(t), bool)
        for (t0, t1) in occl_windows:
            keep &= ~((t >= t0) & (t <= t1))
        t = t[keep]
    pts = observed_curve(t, frame_idx)
    u = P[0, 0] * pts[:, 0] / pts[:, 2] + P[0, 2]; v = P[1, 1] * pts[:, 1] / pts[:, 2] + P[1, 2]
    r = width_px // 2
    zbuf = np.full((rows, cols), np.inf)
    for di in range(-r, r + 1):
        for dj in range(-r, r + 1):
            if di * di + dj * dj > r * r:
                continue
            ii = np.rint(v).astype(int) + di; jj = np.rint(u).astype(int) + dj
            ok = (ii >= 0) & (ii < rows) & (jj >= 0) & (jj < cols)
            np.minimum.at(zbuf, (ii[ok], jj[ok]), pts[ok, 2])
    on = np.isfinite(zbuf)
    col = np.asarray(dlo_bgr, np.int16)
    bgr[on] = np.clip(col + rng.integers(-12, 13, (int(on.sum()), 3)), 0, 255).astype(np.uint8)
    depth[on] = np.rint(zbuf[on] * 1000.0 + rng.normal(0.0, 1.0, int(on.sum()))).astype(np.uint16)
    holes = on & (rng.random((rows, cols)) < depth_holes)
    depth[holes] = 0
    occ = None
    if occlusion_box is not None:
        occ = np.full((rows, cols, 3), 255, np.uint8)
        i0, i1, j0, j1 = occlusion_box
        occ[i0:i1, j0:j1] = 0
    return dict(bgr=bgr, depth=depth, proj=P, occlusion_bgr=occ)
