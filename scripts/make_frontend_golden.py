"""Mints tests/golden/frontend_*.npz: small synthetic camera frames (trackdlo_b200/synth.py: render_frame) with
  * the mask computed by OPENCV ITSELF (cv2.cvtColor BGR2HSV / BGR2GRAY, cv2.inRange, cv2.bitwise_and/or, exactly the calls of
    trackdlo_node.cpp:159-180 and color_thresholding :88-119) -- cv2 is installed in the build container, not on the GPU box;
  * the point cloud of the restatement oracle/frontend.py (back-projection + PCL VoxelGrid semantics; PCL is not installed
    anywhere, see that module's header).
Regenerate with `python scripts/make_frontend_golden.py`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import numpy as np
from oracle import frontend as fe
from trackdlo_b200 import synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def cv2_mask(bgr, lower, upper, multi, occ):
    hsv = cv2.cvtColor(bgr, cv2.COLOR_BGR2HSV)
    if multi:
        m = np.zeros(bgr.shape[:2], np.uint8)
        for lo, hi in fe.MULTI_COLOR_BANDS:
            m = cv2.bitwise_or(m, cv2.inRange(hsv, tuple(float(v) for v in lo), tuple(float(v) for v in hi)))
    else:
        m = cv2.inRange(hsv, tuple(float(v) for v in lower), tuple(float(v) for v in upper))
    if occ is not None:
        m = cv2.bitwise_and(m, cv2.cvtColor(occ, cv2.COLOR_BGR2GRAY))
    return m


CASES = {
    "frontend_blue": dict(idx=0, kw=dict(), multi=False),
    "frontend_occlusion_box": dict(idx=1, kw=dict(occlusion_box=(60, 130, 150, 190)), multi=False),
    "frontend_multi_color": dict(idx=2, kw=dict(dlo_bgr=(30, 200, 230)), multi=True),          # yellow band of color_thresholding
}

for name, c in CASES.items():
    fr = synth.render_frame(c["idx"], rows=180, cols=320, width_px=5, **c["kw"])
    lower, upper = (90, 90, 30), (130, 255, 255)
    m_cv = cv2_mask(fr["bgr"], lower, upper, c["multi"], fr["occlusion_bgr"])
    X, m = fe.point_cloud(fr["bgr"], fr["depth"], fr["proj"], lower, upper, c["multi"], fr["occlusion_bgr"], leaf=0.008)
    assert np.array_equal(m, m_cv), name
    pts = fe.back_project(m, fr["depth"], fr["proj"])
    _, idx, dims = fe.voxel_grid(pts, 0.008)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), bgr=fr["bgr"], depth=fr["depth"], proj=fr["proj"],
                        occlusion_bgr=fr["occlusion_bgr"] if fr["occlusion_bgr"] is not None else np.zeros(0, np.uint8),
                        multi=int(c["multi"]), lower=np.array(lower), upper=np.array(upper), leaf=0.008,
                        cv2_mask=m_cv, cv2_version=cv2.__version__, X=X, voxel_dims=np.array(dims), n_masked=int((m > 0).sum()))
    print(name, "masked", int((m > 0).sum()), "points", len(X), "dims", dims)
