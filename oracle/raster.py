"""TEST INFRASTRUCTURE -- CPU restatement of the self-occlusion test of the visibility block (raster part):
trackdlo/src/trackdlo_node.cpp:280-343

    edges (i, i+1) sorted by the camera distance of their midpoints                       :280-295
    nodes projected through the 3x4 projection matrix, pixel = (int)(x/z), (int)(y/z)       :298-313
    edge by edge, nearest first: a node whose pixel is still 0 in `projected_edges` is not self-occluded (and visible if
    its nearest point is within visibility_threshold); then the edge is drawn with cv::line(.., 255, dlo_pixel_width)   :316-343

Third-party arithmetic that is NOT under /root/reference: OpenCV's cv::line for thickness > 1 (imgproc/src/drawing.cpp:
line -> ThickLine -> FillConvexPoly + Line2 + Circle, clipLine).  opencv-python 4.13 IS installed in the build container, so
this restatement is PINNED: tests/test_raster.py compares `thick_line` with cv2.line pixel for pixel over thousands of random
segments (inside, crossing and outside the image; thickness 1..41), `covers` with the raster, and `self_occlusion` with the
reference's loop run on cv2's own raster; tests/golden/selfocc_*.npz carries cv2's outputs to the GPU box
(scripts/make_selfocc_golden.py).  What cv::line does, as pinned:
  * the segment is first clipped (cv::clipLine, doubles truncated toward zero) to the image rectangle grown by `thickness`
    on every side;
  * ThickLine: the four corners p +- dp, dp = round-half-even(thickness/2 * unit normal) in 16.16 fixed point; the quadrilateral
    is filled by FillConvexPoly (edge walker, x advanced by a rounded 16.16 slope per row) after its outline has been drawn
    with Line2 (a 16.16 DDA, clipped to the image in fixed point); a filled Circle of radius thickness/2 (midpoint algorithm)
    at both ends.
Eigen (also absent) enters through `(proj_matrix * Y_h.transpose())` and `.norm()`: restated as sums in index order."""
import math
import numpy as np

XY_SHIFT = 16
XY_ONE = 1 << XY_SHIFT
HALF = XY_ONE >> 1


def _cdiv(a, b):
    """C integer division (truncation toward zero)."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b > 0) else -q


def clip_line(w, h, x1, y1, x2, y2):
    """cv::clipLine(Size2l(w, h), pt1, pt2) -> (visible, x1, y1, x2, y2)."""
    right, bottom = w - 1, h - 1
    if w <= 0 or h <= 0:
        return False, x1, y1, x2, y2
    c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8
    c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8
    if (c1 & c2) == 0 and (c1 | c2) != 0:
        if c1 & 12:
            a = 0 if c1 < 8 else bottom
            x1 += int(float(a - y1) * float(x2 - x1) / float(y2 - y1)); y1 = a
            c1 = (x1 < 0) + (x1 > right) * 2
        if c2 & 12:
            a = 0 if c2 < 8 else bottom
            x2 += int(float(a - y2) * float(x2 - x1) / float(y2 - y1)); y2 = a
            c2 = (x2 < 0) + (x2 > right) * 2
        if (c1 & c2) == 0 and (c1 | c2) != 0:
            if c1:
                a = 0 if c1 == 1 else right
                y1 += int(float(a - x1) * float(y2 - y1) / float(x2 - x1)); x1 = a; c1 = 0
            if c2:
                a = 0 if c2 == 1 else right
                y2 += int(float(a - x2) * float(y2 - y1) / float(x2 - x1)); x2 = a; c2 = 0
    return (c1 | c2) == 0, x1, y1, x2, y2


def _line2_points(W, H, p1, p2):
    """Line2: the pixels of the 16.16 DDA between two fixed-point points (generator of (x, y), unclipped per pixel)."""
    ok, x1, y1, x2, y2 = clip_line(W << XY_SHIFT, H << XY_SHIFT, p1[0], p1[1], p2[0], p2[1])
    if not ok:
        return
    dx, dy = x2 - x1, y2 - y1
    j = -1 if dx < 0 else 0
    ax = (dx ^ j) - j
    i = -1 if dy < 0 else 0
    ay = (dy ^ i) - i
    if ax > ay:
        dy = (dy ^ j) - j
        if j:
            x1, x2, y1, y2 = x2, x1, y2, y1
        x_step, y_step = XY_ONE, _cdiv(dy * XY_ONE, ax | 1)
        ecount = (x2 - x1) >> XY_SHIFT
    else:
        dx = (dx ^ i) - i
        if i:
            x1, x2, y1, y2 = x2, x1, y2, y1
        x_step, y_step = _cdiv(dx * XY_ONE, ay | 1), XY_ONE
        ecount = (y2 - y1) >> XY_SHIFT
    x1 += HALF; y1 += HALF
    yield (x2 + HALF) >> XY_SHIFT, (y2 + HALF) >> XY_SHIFT
    while ecount >= 0:
        yield x1 >> XY_SHIFT, y1 >> XY_SHIFT
        x1 += x_step; y1 += y_step; ecount -= 1


def _fill_rows(W, H, v):
    """FillConvexPoly's scan conversion (without the outline): generator of (y, x_first, x_last), unclipped in x."""
    npts = len(v)
    delta = HALF
    ymin = ymax = v[0][1]; xmin = xmax = v[0][0]; imin = 0
    for i, p in enumerate(v):
        if p[1] < ymin:
            ymin = p[1]; imin = i
        ymax = max(ymax, p[1]); xmax = max(xmax, p[0]); xmin = min(xmin, p[0])
    xmin = (xmin + delta) >> XY_SHIFT; xmax = (xmax + delta) >> XY_SHIFT
    ymin = (ymin + delta) >> XY_SHIFT; ymax = (ymax + delta) >> XY_SHIFT
    if xmax < 0 or ymax < 0 or xmin >= W or ymin >= H:
        return
    ymax = min(ymax, H - 1)
    e = [dict(idx=imin, di=1, x=-XY_ONE, dx=0, ye=ymin), dict(idx=imin, di=npts - 1, x=-XY_ONE, dx=0, ye=ymin)]
    y, edges = ymin, npts
    while True:
        for i in range(2):
            if y >= e[i]["ye"]:
                idx0, di = e[i]["idx"], e[i]["di"]
                idx = idx0 + di
                if idx >= npts:
                    idx -= npts
                while True:
                    edges -= 1
                    if edges < 0:
                        break
                    ty = (v[idx][1] + delta) >> XY_SHIFT
                    if ty > y:
                        xs, xe = v[idx0][0], v[idx][0]
                        e[i].update(ye=ty, dx=_cdiv((xe - xs) * 2 + (ty - y), 2 * (ty - y)), x=xs, idx=idx)
                        break
                    idx0 = idx; idx += di
                    if idx >= npts:
                        idx -= npts
        if edges < 0:
            break
        if y >= 0:
            l, r = (0, 1) if e[0]["x"] <= e[1]["x"] else (1, 0)
            xx1 = (e[l]["x"] + HALF) >> XY_SHIFT; xx2 = (e[r]["x"] + HALF) >> XY_SHIFT
            if xx2 >= 0 and xx1 < W:
                yield y, xx1, xx2
        e[0]["x"] += e[0]["dx"]; e[1]["x"] += e[1]["dx"]
        y += 1
        if y > ymax:
            break


def circle_halfwidths(radius):
    """Filled cv::Circle: half-width of the span at row offset k = 0..radius (-1: no pixel)."""
    hw = [-1] * (radius + 1)
    err, dx, dy, plus, minus = 0, radius, 0, 1, (radius << 1) - 1
    while dx >= dy:
        hw[dy] = max(hw[dy], dx)
        hw[dx] = max(hw[dx], dy)
        dy += 1; err += plus; plus += 2
        mask = -1 if err > 0 else 0
        err -= minus & mask; dx += mask; minus -= mask & 2
    return hw


def _thick_geometry(W, H, a, b, thickness):
    """Pre-clip + ThickLine set-up: (quad corners in 16.16 or None, [circle centres], radius) or None when nothing is drawn."""
    m = thickness
    ok, x1, y1, x2, y2 = clip_line(W + 2 * m, H + 2 * m, a[0] + m, a[1] + m, b[0] + m, b[1] + m)
    if not ok:
        return None
    p0 = ((x1 - m) << XY_SHIFT, (y1 - m) << XY_SHIFT); p1 = ((x2 - m) << XY_SHIFT, (y2 - m) << XY_SHIFT)
    inv = 1.0 / XY_ONE
    dx = (p0[0] - p1[0]) * inv; dy = (p1[1] - p0[1]) * inv
    r = dx * dx + dy * dy
    odd = thickness & 1
    th = thickness << (XY_SHIFT - 1)
    quad = None
    if abs(r) > 2.220446049250313e-16:
        r = (th + odd * XY_ONE * 0.5) / math.sqrt(r)
        dpx, dpy = int(np.rint(dy * r)), int(np.rint(dx * r))
        quad = [(p0[0] + dpx, p0[1] + dpy), (p0[0] - dpx, p0[1] - dpy), (p1[0] - dpx, p1[1] - dpy), (p1[0] + dpx, p1[1] + dpy)]
    centres = [((p[0] + HALF) >> XY_SHIFT, (p[1] + HALF) >> XY_SHIFT) for p in (p0, p1)]
    return quad, centres, (th + HALF) >> XY_SHIFT


def thick_line(img, a, b, thickness):
    """cv::line(img, a, b, 255, thickness) for thickness > 1, LINE_8, shift 0, 8-bit single channel: rasterised into img."""
    H, W = img.shape
    g = _thick_geometry(W, H, a, b, thickness)
    if g is None:
        return
    quad, centres, radius = g

    def hline(y, xa, xb):
        if 0 <= y < H:
            xa, xb = max(xa, 0), min(xb, W - 1)
            if xa <= xb:
                img[y, xa:xb + 1] = 255
    if quad is not None:
        for i in range(4):
            for x, y in _line2_points(W, H, quad[i - 1], quad[i]):
                if 0 <= x < W and 0 <= y < H:
                    img[y, x] = 255
        for y, xa, xb in _fill_rows(W, H, quad):
            hline(y, xa, xb)
    hw = circle_halfwidths(radius)
    for cx, cy in centres:
        for k, w in enumerate(hw):
            if w >= 0:
                hline(cy - k, cx - w, cx + w); hline(cy + k, cx - w, cx + w)


def covers(px, py, W, H, a, b, thickness):
    """True iff cv::line(img, a, b, 255, thickness) sets pixel (px, py) of a W x H image -- without a raster (this is the form
    the device code evaluates per (node, edge) pair)."""
    if not (0 <= px < W and 0 <= py < H):
        return False
    g = _thick_geometry(W, H, a, b, thickness)
    if g is None:
        return False
    quad, centres, radius = g
    hw = circle_halfwidths(radius)
    for cx, cy in centres:
        k = abs(py - cy)
        if k <= radius and hw[k] >= 0 and abs(px - cx) <= hw[k]:
            return True
    if quad is None:
        return False
    for y, xa, xb in _fill_rows(W, H, quad):
        if y == py:
            if xa <= px <= xb:
                return True
            break
        if y > py:
            break
    for i in range(4):
        for x, y in _line2_points(W, H, quad[i - 1], quad[i]):
            if x == px and y == py:
                return True
    return False


def project_pixels(Y, proj):
    """trackdlo_node.cpp:298-313: (col, row) of every node; the matrix product restated as a sum in index order."""
    out = []
    for p in np.asarray(Y, float):
        h = (p[0], p[1], p[2], 1.0)
        ic = []
        for r in range(3):
            s = 0.0
            for k in range(4):
                s = s + proj[r, k] * h[k]
            ic.append(s)
        out.append((int(ic[0] / ic[2]), int(ic[1] / ic[2])))     # static_cast<int>: truncation toward zero
    return out


def edge_order(Y):
    """trackdlo_node.cpp:280-295: edge indices sorted by the norm of the edge midpoint (ties: by index; std::sort leaves them
    unspecified)."""
    Y = np.asarray(Y, float)
    d = []
    for i in range(len(Y) - 1):
        mx, my, mz = (Y[i, 0] + Y[i + 1, 0]) / 2, (Y[i, 1] + Y[i + 1, 1]) / 2, (Y[i, 2] + Y[i + 1, 2]) / 2
        d.append(math.sqrt((mx * mx + my * my) + mz * mz))
    return sorted(range(len(d)), key=lambda i: (d[i], i))


def self_occlusion(Y, proj, rows, cols, pixel_width, use_raster=False, line_fn=None):
    """trackdlo_node.cpp:316-343.  Returns not_self_occluded [Nn] bool.  A node is tested at its first visit only (later visits
    see a superset of the lines).  Pixels outside the image read as 0 (the reference indexes out of bounds there).
    use_raster=True draws into an image like the reference (line_fn = cv2.line in the pinning tests)."""
    n = len(Y)
    pix = project_pixels(Y, proj)
    order = edge_order(Y)
    seen = [False] * n
    free = [False] * n
    if use_raster:
        img = np.zeros((rows, cols), np.uint8)
        for e in order:
            for node in (e, e + 1):
                c, r = pix[node]
                if 0 <= r < rows and 0 <= c < cols:
                    if img[r, c] == 0:
                        free[node] = True
                else:
                    free[node] = True
            if line_fn is None:
                thick_line(img, pix[e], pix[e + 1], pixel_width)
            else:
                line_fn(img, pix[e], pix[e + 1], 255, pixel_width)
        return np.array(free)
    drawn = []
    for e in order:
        for node in (e, e + 1):
            if seen[node]:
                continue
            seen[node] = True
            c, r = pix[node]
            free[node] = not any(covers(c, r, cols, rows, pix[d], pix[d + 1], pixel_width) for d in drawn)
        drawn.append(e)
    return np.array(free)


def visible_nodes(Y, proj, rows, cols, pixel_width, dmin, visibility_threshold):
    """visible_nodes of trackdlo_node.cpp:316-346 (sorted): not self-occluded and within visibility_threshold of a point."""
    free = self_occlusion(Y, proj, rows, cols, pixel_width)
    return np.array([i for i in range(len(Y)) if free[i] and dmin[i] <= visibility_threshold], np.int32)
