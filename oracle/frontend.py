"""TEST INFRASTRUCTURE -- CPU restatement (NumPy) of the perception front-end that produces the tracker's input cloud
(SURVEY.md §8 f2): trackdlo/src/trackdlo_node.cpp:159-242

    BGR image -> HSV (cv::cvtColor COLOR_BGR2HSV, 8-bit)                    :159
    -> cv::inRange per colour band / color_thresholding                     :161-167, :88-119
    -> AND with the grey occlusion mask (cv::COLOR_BGR2GRAY, bitwise_and)   :172-180
    -> mask + depth (uint16 mm) -> points through the projection matrix      :195-233
    -> pcl::VoxelGrid centroid down-sampling, leaf 0.008                     :236-240
    -> float32 points widened to double = X                                  :242

Third-party arithmetic that is NOT under /root/reference:
  * OpenCV (imgproc): the 8-bit BGR2HSV / BGR2GRAY fixed-point formulas and inRange.  opencv-python 4.13 IS installed in
    the build container, so those three are PINNED: tests/test_frontend.py compares them with cv2 itself over all 2^24
    colours, and tests/golden/frontend_*.npz carries cv2's outputs to the GPU box (scripts/make_frontend_golden.py).
  * PCL 1.10 (ROS noetic) pcl::VoxelGrid<PointXYZRGB> (pcl/filters/impl/voxel_grid.hpp): NOT installed anywhere here ->
    restated from the published algorithm, PARITY UNPINNED for this stage: voxel index of a point =
    floor(p * (1/leaf)) in float32, relative to the floor of the cloud's minimum; one output point per occupied voxel,
    in ascending order of the linear index i + j*dx + k*dx*dy; output point = centroid of the voxel's points.
    Where this restatement departs from PCL: PCL accumulates the centroid in float32 in the order its (unstable)
    std::sort leaves the points, so its last bits depend on libstdc++'s introsort; here the centroid is the exact mean
    (order-independent: sums of the coordinates in fixed point, 2^-36 m, exact for float32 inputs >= 0.25 mm) rounded
    once to float32.  Voxel membership, the number of output points and their order are the same; coordinates can
    differ from PCL's by a few float32 ulps (~1e-7 relative)."""
import numpy as np

HSV_SHIFT = 12
FIX = 68719476736.0          # 2^36


def _div_tables():
    # OpenCV RGB2HSV_b: sdiv_table[i] = saturate_cast<int>((255 << hsv_shift) / (1.*i)), hdiv_table[i] = ((180 << hsv_shift) / (6.*i))
    sdiv = np.zeros(256, np.int64); hdiv = np.zeros(256, np.int64)
    i = np.arange(1, 256, dtype=np.float64)
    sdiv[1:] = np.rint((255 << HSV_SHIFT) / i).astype(np.int64)
    hdiv[1:] = np.rint((180 << HSV_SHIFT) / (6.0 * i)).astype(np.int64)
    return sdiv, hdiv


_SDIV, _HDIV = _div_tables()


def bgr2hsv(bgr):
    """cv::cvtColor(..., COLOR_BGR2HSV) for 8-bit images: H in [0,180), S, V in [0,255]."""
    b = bgr[..., 0].astype(np.int64); g = bgr[..., 1].astype(np.int64); r = bgr[..., 2].astype(np.int64)
    v = np.maximum(np.maximum(b, g), r); vmin = np.minimum(np.minimum(b, g), r)
    diff = v - vmin
    s = (diff * _SDIV[v] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT
    h = np.where(v == r, g - b, np.where(v == g, b - r + 2 * diff, r - g + 4 * diff))
    h = (h * _HDIV[diff] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT
    h = h + np.where(h < 0, 180, 0)
    return np.stack([h, s, v], axis=-1).astype(np.uint8)


def bgr2gray(bgr):
    """cv::cvtColor(..., COLOR_BGR2GRAY), 8-bit, as OpenCV 4.13 computes it: (B*3735 + G*19235 + R*9798 + 16384) >> 15.
    (OpenCV <= 4.1 used the 14-bit coefficients 1868 / 9617 / 4899; the two formulas give 0 for exactly the same colours,
    checked over all 2^24 -- and `mask & grey` (trackdlo_node.cpp:175) only asks whether grey is 0.)"""
    b = bgr[..., 0].astype(np.int64); g = bgr[..., 1].astype(np.int64); r = bgr[..., 2].astype(np.int64)
    return ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)


def in_range(hsv, lower, upper):
    lo = np.asarray(lower).reshape(1, 1, 3); hi = np.asarray(upper).reshape(1, 1, 3)
    h = hsv.astype(np.int64)
    return (np.all((h >= lo) & (h <= hi), axis=-1) * 255).astype(np.uint8)


# color_thresholding (trackdlo_node.cpp:88-119): blue | red_1 | red_2 | yellow
MULTI_COLOR_BANDS = (((90, 90, 60), (130, 255, 255)), ((130, 60, 50), (255, 255, 255)), ((0, 60, 50), (10, 255, 255)),
                     ((15, 100, 80), (40, 255, 255)))


def dlo_mask(bgr, lower, upper, multi_color=False, occlusion_bgr=None):
    hsv = bgr2hsv(bgr)
    if multi_color:
        m = np.zeros(bgr.shape[:2], np.uint8)
        for lo, hi in MULTI_COLOR_BANDS:
            m |= in_range(hsv, lo, hi)
    else:
        m = in_range(hsv, lower, upper)
    if occlusion_bgr is not None:
        m = m & bgr2gray(occlusion_bgr)              # bitwise, as cv::bitwise_and
    return m


def back_project(mask, depth, proj):
    """Masked pixels in row-major order -> float32 points (the fields of pcl::PointXYZRGB), trackdlo_node.cpp:210-222."""
    ii, jj = np.nonzero(mask)
    fx, fy, cx, cy = proj[0, 0], proj[1, 1], proj[0, 2], proj[1, 2]
    z = depth[ii, jj].astype(np.float64) / 1000.0
    x = (jj.astype(np.float64) - cx) * z / fx
    y = (ii.astype(np.float64) - cy) * z / fy
    return np.stack([x, y, z], axis=1).astype(np.float32)


def voxel_grid(points_f32, leaf):
    """pcl::VoxelGrid (see the module docstring).  Returns (centroids float32 [n,3], linear voxel index per input point,
    dims).  dims is None and the input is returned unchanged when the grid would overflow an int (PCL's own bail-out)."""
    p = np.asarray(points_f32, np.float32)
    if len(p) == 0:
        return p.reshape(0, 3), np.zeros(0, np.int64), (0, 0, 0)
    inv = np.float32(1.0) / np.float32(leaf)
    ijk = np.floor(p * inv).astype(np.int64)                      # float32 product, like PCL
    mn = ijk.min(axis=0); mx = ijk.max(axis=0)
    dims = mx - mn + 1
    if int(dims[0]) * int(dims[1]) * int(dims[2]) > np.iinfo(np.int32).max:
        return p, np.arange(len(p)), None
    rel = ijk - mn
    idx = rel[:, 0] + rel[:, 1] * dims[0] + rel[:, 2] * dims[0] * dims[1]
    uniq, inv_idx, cnt = np.unique(idx, return_inverse=True, return_counts=True)      # ascending index = PCL's output order
    # order-independent mean: integer sums of the coordinates in units of 2^-36 m
    q = np.rint(p.astype(np.float64) * FIX).astype(np.int64)
    sums = np.zeros((len(uniq), 3), np.int64)
    np.add.at(sums, inv_idx, q)
    out = (sums.astype(np.float64) / cnt[:, None].astype(np.float64)) * (1.0 / FIX)
    return out.astype(np.float32), idx, tuple(int(v) for v in dims)


def point_cloud(bgr, depth, proj, lower=(90, 90, 30), upper=(130, 255, 255), multi_color=False, occlusion_bgr=None, leaf=0.008):
    """The whole front-end for one frame: X [Mp,3] float64 (float32-valued), plus the intermediate mask."""
    m = dlo_mask(bgr, lower, upper, multi_color, occlusion_bgr)
    pts = back_project(m, depth, proj)
    cen, _, _ = voxel_grid(pts, leaf)
    return cen.astype(np.float64), m
