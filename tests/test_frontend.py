"""Perception front-end (trackdlo_node.cpp:159-242).

CPU (`-m "not gpu"`): the restatement oracle/frontend.py against OpenCV itself where OpenCV is importable (the build
container) and against the cv2-minted goldens everywhere; known answers for the PCL VoxelGrid semantics.
GPU (`-m gpu`): tdlo_point_cloud_batched against the goldens / the oracle -- mask-derived integers (point counts, voxel
membership via the order of the output) exact, coordinates bit-exact -- and the chain camera frame -> cloud -> visibility
-> tracking_step on the device against the oracle chain."""
import glob
import os

import numpy as np
import pytest

import oracle
from oracle import frontend as fe
from trackdlo_b200 import api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDENS = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "frontend_*.npz")))

try:
    import cv2
except Exception:           # not on the GPU box
    cv2 = None


def _load(path):
    g = np.load(path)
    occ = g["occlusion_bgr"] if g["occlusion_bgr"].size else None
    return g, occ


# ------------------------------------------------------------------------------------------------ CPU
@pytest.mark.skipif(cv2 is None, reason="cv2 not importable here")
def test_hsv_grey_inrange_equal_opencv_on_all_colours():
    v = np.arange(1 << 24, dtype=np.uint32)
    img = np.stack([v & 255, (v >> 8) & 255, (v >> 16) & 255], axis=-1).astype(np.uint8).reshape(4096, 4096, 3)
    hsv = cv2.cvtColor(img, cv2.COLOR_BGR2HSV)
    mine = fe.bgr2hsv(img)
    assert np.array_equal(hsv, mine)
    assert np.array_equal(cv2.cvtColor(img, cv2.COLOR_BGR2GRAY), fe.bgr2gray(img))
    for lo, hi in (((90, 90, 30), (130, 255, 255)),) + fe.MULTI_COLOR_BANDS:
        assert np.array_equal(cv2.inRange(hsv, tuple(float(x) for x in lo), tuple(float(x) for x in hi)), fe.in_range(mine, lo, hi))


@pytest.mark.parametrize("path", GOLDENS, ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_mask_equals_opencv_golden(path):
    g, occ = _load(path)
    m = fe.dlo_mask(g["bgr"], tuple(g["lower"]), tuple(g["upper"]), bool(g["multi"]), occ)
    assert np.array_equal(m, g["cv2_mask"])
    X, _ = fe.point_cloud(g["bgr"], g["depth"], g["proj"], tuple(g["lower"]), tuple(g["upper"]), bool(g["multi"]), occ, float(g["leaf"]))
    assert np.array_equal(X, g["X"])


def test_voxel_grid_known_answers():
    leaf = 0.008
    # three points in one voxel, one in the next along x, one far along z; a point with negative coordinates moves the origin
    p = np.array([[0.0161, 0.0001, 0.0001], [0.0165, 0.0003, 0.0002], [0.0170, 0.0002, 0.0079], [0.0241, 0.0, 0.0], [0.0161, 0.0, 0.0801],
                  [-0.0001, -0.0001, -0.0001]], np.float32)
    cen, idx, dims = fe.voxel_grid(p, leaf)
    assert dims == (5, 2, 12)                                   # floor(p / leaf): x in -1..3, y in -1..0, z in -1..10
    assert list(idx) == [18, 18, 18, 19, 118, 0]                # i + j*dx + k*dx*dy relative to the minimum (-1, -1, -1): (3,1,1) -> 3 + 5 + 10
    assert len(cen) == 4 and np.allclose(cen[0], p[5]) and np.allclose(cen[1], p[:3].mean(axis=0), atol=1e-7)
    assert np.allclose(cen[2], p[3]) and np.allclose(cen[3], p[4])
    # one point per occupied voxel, sorted by linear index; leaf boundaries: x = 0.008 belongs to voxel 1 (floor)
    q = np.array([[0.0079999, 0, 0], [0.008, 0, 0], [0.0160001, 0, 0]], np.float32)
    cen, idx, dims = fe.voxel_grid(q, leaf)
    assert dims == (3, 1, 1) and list(idx) == [0, 1, 2]
    # empty input, and the int-overflow bail-out of PCL
    assert fe.voxel_grid(np.zeros((0, 3), np.float32), leaf)[0].shape == (0, 3)
    far = np.array([[0, 0, 0], [30.0, 30.0, 30.0]], np.float32)
    assert fe.voxel_grid(far, leaf)[2] is None


def test_back_projection_keeps_zero_depth_pixels_and_casts_to_float32():
    P = synth.camera_matrix(4, 6, f=500.0)
    mask = np.zeros((4, 6), np.uint8); mask[1, 2] = 255; mask[3, 5] = 255
    depth = np.zeros((4, 6), np.uint16); depth[1, 2] = 650
    pts = fe.back_project(mask, depth, P)
    assert pts.dtype == np.float32 and pts.shape == (2, 3)
    assert np.array_equal(pts[1], [0, 0, 0])                   # depth 0 is NOT filtered by the reference (trackdlo_node.cpp:216)
    assert pts[0, 2] == np.float32(0.65) and pts[0, 0] == np.float32((2 - P[0, 2]) * 0.65 / 500.0)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDENS, ids=lambda p: os.path.basename(p)[:-4])
def test_gpu_point_cloud_equals_golden(path):
    g, occ = _load(path)
    F = 3                                                        # the same frame three times + a blank one in between
    blank = np.full_like(g["bgr"], 127)
    bgr = np.stack([g["bgr"], blank, g["bgr"]]); depth = np.stack([g["depth"]] * 3); proj = np.stack([g["proj"]] * 3)
    occs = None if occ is None else np.stack([occ] * 3)
    ctx = api.Context(max_frames=F, max_nodes=30, max_points_total=20000)
    try:
        r = ctx.point_cloud_batched(bgr, depth, proj, tuple(int(v) for v in g["lower"]), tuple(int(v) for v in g["upper"]), bool(g["multi"]), occs, float(g["leaf"]))
        n = len(g["X"])
        assert list(r["x_offsets"]) == [0, n, n, 2 * n]
        assert list(r["status"]) == [0, api.FE_EMPTY, 0]
        assert np.array_equal(r["X"][:n], g["X"]) and np.array_equal(r["X"][n:], g["X"])       # bit-exact, order included
    finally:
        ctx.close()


@pytest.mark.gpu
def test_gpu_point_cloud_full_resolution_and_capacity():
    fr = [synth.render_frame(i, occlusion_box=(300, 420, 600, 700) if i == 1 else None) for i in range(2)]
    occ = np.stack([f["occlusion_bgr"] if f["occlusion_bgr"] is not None else np.full_like(f["bgr"], 255) for f in fr])
    bgr = np.stack([f["bgr"] for f in fr]); depth = np.stack([f["depth"] for f in fr]); proj = np.stack([f["proj"] for f in fr])
    want = [fe.point_cloud(f["bgr"], f["depth"], f["proj"], occlusion_bgr=o)[0] for f, o in zip(fr, occ)]
    ctx = api.Context(max_frames=2, max_nodes=30, max_points_total=5000)
    try:
        r = ctx.point_cloud_batched(bgr, depth, proj, occlusion_bgr=occ)
        assert list(r["x_offsets"]) == [0, len(want[0]), len(want[0]) + len(want[1])] and list(r["status"]) == [0, 0]
        assert np.array_equal(r["X"], np.concatenate(want))
        # X too small for the second frame: it (and nothing else) is dropped and flagged
        r2 = ctx.point_cloud_batched(bgr, depth, proj, occlusion_bgr=occ, x_capacity=len(want[0]) + 3)
        assert list(r2["x_offsets"]) == [0, len(want[0]), len(want[0])] and list(r2["status"]) == [0, api.FE_CAPACITY]
        assert np.array_equal(r2["X"], want[0])
        # a grid workspace too small for the frame: refused with TDLO_FE_GRID
        ctx.set_option("voxel_cells", 4096)
        r3 = ctx.point_cloud_batched(bgr[:1], depth[:1], proj[:1])
        assert list(r3["status"]) == [api.FE_GRID] and r3["x_offsets"][1] == 0
    finally:
        ctx.close()


@pytest.mark.gpu
def test_camera_frame_to_nodes_on_the_device():
    """Depth image in -> nodes out, nothing but device pointers in between: tdlo_point_cloud_batched_device ->
    tdlo_visibility_batched_device -> tdlo_tracking_step_batched_device, for a short sequence of frames; against the oracle
    chain (oracle/frontend.py -> oracle.visibility -> oracle.tracking_step) driven the same way."""
    import ctypes as C
    import torch
    dev = torch.device("cuda:0")
    N, T = 45, 3
    Y0 = synth.curve(np.linspace(0, 1, N)); rest = synth.rest_arclengths(Y0)
    frames = [synth.render_frame(t, occl_windows=[(0.4, 0.55)] if t == 2 else None) for t in range(T)]
    cap = 4000
    ctx = api.Context(max_frames=1, max_nodes=N, max_points_total=cap)
    try:
        t_ = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        dY = t_(Y0[None].copy()); ds2 = torch.zeros(1, dtype=torch.float64, device=dev); drest = t_(rest[None])
        dX = torch.zeros(cap, 3, dtype=torch.float64, device=dev); dxo = torch.zeros(2, dtype=torch.int64, device=dev)
        dvis = torch.zeros(N, dtype=torch.int32, device=dev); dext = torch.zeros(N, dtype=torch.int32, device=dev)
        dvo = torch.zeros(2, dtype=torch.int64, device=dev); deo = torch.zeros(2, dtype=torch.int64, device=dev)
        dit = torch.zeros(1, 2, dtype=torch.int32, device=dev); dst = torch.zeros(1, dtype=torch.int32, device=dev)
        stream = torch.cuda.current_stream()
        Yo, s2o = Y0.copy(), 0.0
        tp, otp = api.TrackParams(), oracle.TrackParams()
        for t in range(T):
            fr = frames[t]
            dbgr, ddep, dproj = t_(fr["bgr"][None]), t_(fr["depth"][None]), t_(fr["proj"].reshape(1, 12))
            fb = api.FrontendBatchC(1, fr["bgr"].shape[0], fr["bgr"].shape[1], 0, dbgr.data_ptr(), ddep.data_ptr(), None, dproj.data_ptr(),
                                    (C.c_int32 * 3)(90, 90, 30), (C.c_int32 * 3)(130, 255, 255), 0.008, dX.data_ptr(), dxo.data_ptr(), cap, None)
            ctx.point_cloud_batched_raw(fb, device=True, stream=stream.cuda_stream)
            vb = api.VisBatchC(1, N, dX.data_ptr(), dxo.data_ptr(), dY.data_ptr(), drest.data_ptr(), tp.visibility_threshold, 0.06, None,
                               dvis.data_ptr(), dvo.data_ptr(), dext.data_ptr(), deo.data_ptr())
            ctx.visibility_batched_raw(vb, device=True, stream=stream.cuda_stream)
            tb = api.TrackBatchC(1, N, dX.data_ptr(), dxo.data_ptr(), dY.data_ptr(), ds2.data_ptr(), drest.data_ptr(), dvis.data_ptr(), dvo.data_ptr(),
                                 dext.data_ptr(), deo.data_ptr(), None, None, None, None, dit.data_ptr(), dst.data_ptr(), None)
            ctx.tracking_step_batched_raw(tb, tp.to_c(), device=True, stream=stream.cuda_stream)
            ctx.synchronize()
            # oracle chain, from the SAME carried state (the pre-processing registration's LLE weights amplify 1e-13 differences)
            X, _ = fe.point_cloud(fr["bgr"], fr["depth"], fr["proj"])
            v = oracle.visibility(X, Yo, rest, otp.visibility_threshold, 0.06)
            o = oracle.tracking_step(X, Yo, s2o, rest, v["vis"], v["vis_ext"], otp)
            assert int(dxo[1]) == len(X) and np.array_equal(dX[:len(X)].cpu().numpy(), X)
            assert list(dit.cpu().numpy()[0]) == list(o["iters"]) and int(dst[0]) & ~(api.ST_NOT_CONVERGED | api.ST_PRE_NOT_CONVERGED) == 0
            Yg = dY.cpu().numpy()[0]
            assert float(np.abs(Yg - o["Y"]).max() / np.abs(o["Y"]).max()) < 1e-6, t
            Yo, s2o = Yg.copy(), float(ds2[0])                 # carry the device state on both sides
            dY.copy_(t_(Yo[None]))
        truth = synth.observed_curve(np.linspace(0, 1, N), T - 1)
        assert np.abs(Yo - truth).max() < 0.03                 # and it actually tracks the rope
    finally:
        ctx.close()
