// Visibility front-end feeding tracking_step, batched over independent frames:
//   shortest node-to-point distances           trackdlo/src/trackdlo_node.cpp:254-277
//   visible_nodes (sorted) / visible_nodes_extended (d_vis rule)   trackdlo_node.cpp:346-360
//   self-occlusion test (optional)               trackdlo_node.cpp:280-343  (tdlo_vis_selfocc_kernel below)
// Small kernels, no host round trip (the device-pointer entry point is fully stream-ordered):
// (0) one CTA: slice table = exclusive scan of ceil(Mp_f / VIS_SLICE) over the frames; (1) per-node min squared
// distance, a fixed grid of CTAs striding over the (frame, slice) pairs, combined with atomicMin on the bit pattern
// (non-negative doubles order like their bits); (2) per frame: sqrt, threshold, the two lists at a fixed stride +
// their lengths; (3) one CTA: exclusive scan of the lengths -> CSR offsets, compaction.  The CSR outputs are exactly
// the visibility inputs of tdlo_tracking_step_batched_device.
#pragma once

#include "tdlo_common.cuh"

namespace tdlo {

struct VisArgs {
    int n_frames, n_nodes;
    const double* X; const long long* x_off;
    const double* Y; const double* node_coord;
    double tau, d_vis;
    unsigned long long* dmin2_bits;   // [F][N] workspace
    int* tmp_vis; int* tmp_ext;       // [F][N] workspace
    int* counts;                      // [F][2] workspace
    double* dmin_out;                 // optional [F][N]
    long long* slice_start;           // [F+1] workspace: first global slice of every frame
    long long max_points;             // capacity: frames whose offsets exceed it contribute no slices
    int* vis; long long* vis_off; int* ext; long long* ext_off;
    // self-occlusion test (proj == nullptr: off, every node counts as not self-occluded)
    const double* proj;               // [F][12] row-major 3x4 projection matrices
    int rows, cols, pixel_width;      // image size, dlo_pixel_width (cv::line thickness, >= 2)
    int* free_flag;                   // [F][N] workspace / optional output: 1 = not self-occluded
};

constexpr int VIS_SLICE = 4096;       // points per CTA of kernel 1

__global__ void __launch_bounds__(256) tdlo_vis_slices_kernel(const VisArgs a) {
    __shared__ long long part[256];
    __shared__ long long run;
    const int tid = threadIdx.x, F = a.n_frames;
    if (tid == 0) { run = 0; a.slice_start[0] = 0; }
    __syncthreads();
    for (int f0 = 0; f0 < F; f0 += 256) {
        const int f = f0 + tid;
        long long ns = 0;
        if (f < F) {
            const long long x0 = a.x_off[f], x1 = a.x_off[f + 1];
            if (x0 >= 0 && x1 >= x0 && x1 <= a.max_points) ns = (x1 - x0 + VIS_SLICE - 1) / VIS_SLICE;
        }
        part[tid] = ns;
        __syncthreads();
        if (tid == 0) { long long r = run; for (int i = 0; i < 256; i++) { const long long v = part[i]; part[i] = r; r += v; } run = r; }
        __syncthreads();
        if (f < F) a.slice_start[f + 1] = part[tid] + ns;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) tdlo_vis_dmin_kernel(const VisArgs a) {
    __shared__ double4 nd[kMaxNodes];
    __shared__ int s_f;
    const int N = a.n_nodes, tid = threadIdx.x, lane = tid & 31, F = a.n_frames;
    const long long total = a.slice_start[F];
    for (long long sl = blockIdx.x; sl < total; sl += gridDim.x) {
        if (tid == 0) {                                            // frame of this slice: last f with slice_start[f] <= sl
            int lo = 0, hi = F - 1;
            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (a.slice_start[mid] <= sl) lo = mid; else hi = mid - 1; }
            s_f = lo;
        }
        __syncthreads();
        const int f = s_f;
        const long long x0 = a.x_off[f], m0 = a.x_off[f + 1] - x0;
        const long long p0 = (sl - a.slice_start[f]) * VIS_SLICE;
        const long long p1 = p0 + VIS_SLICE < m0 ? p0 + VIS_SLICE : m0;
        for (int j = tid; j < N; j += blockDim.x) nd[j] = make_double4(a.Y[((long long)f * N + j) * 3], a.Y[((long long)f * N + j) * 3 + 1], a.Y[((long long)f * N + j) * 3 + 2], 0.0);
        __syncthreads();
        const double* X = a.X + x0 * 3;
        for (long long base = p0 + (tid & ~31); base < p1; base += blockDim.x) {
            const long long n = base + lane;
            const bool valid = n < p1;
            double x = 0, y = 0, z = 0;
            if (valid) { x = __ldg(X + n * 3); y = __ldg(X + n * 3 + 1); z = __ldg(X + n * 3 + 2); }
            for (int j = 0; j < N; j++) {
                const double4 q = nd[j];
                const double dx = q.x - x, dy = q.y - y, dz = q.z - z;
                // same operation order as the reference's (Y.row(m) - X.row(n)).norm(), no fused multiply-add
                double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                if (!valid) d2 = 1e300;
                const unsigned hi = (unsigned)__double2hiint(d2);
                const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
                const unsigned lw = (hi == mh) ? (unsigned)__double2loint(d2) : 0xffffffffu;
                const unsigned ml = __reduce_min_sync(0xffffffffu, lw);
                if (lane == 0) atomicMin(a.dmin2_bits + (long long)f * N + j, ((unsigned long long)mh << 32) | ml);
            }
        }
        __syncthreads();
    }
}

__global__ void tdlo_vis_lists_kernel(const VisArgs a) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= a.n_frames) return;
    const int N = a.n_nodes;
    int* vis = a.tmp_vis + (long long)f * N;
    int* ext = a.tmp_ext + (long long)f * N;
    const double* nc = a.node_coord + (long long)f * N;
    int nv = 0;
    for (int m = 0; m < N; m++) {
        const double d2 = __longlong_as_double((long long)a.dmin2_bits[(long long)f * N + m]);
        double d = sqrt(d2);
        if (!(d < 100000.0)) d = 100000.0;                        // the reference's initial shortest_dist
        if (a.dmin_out) a.dmin_out[(long long)f * N + m] = d;
        if (d <= a.tau && (!a.proj || a.free_flag[(long long)f * N + m])) vis[nv++] = m;
    }
    int ne = 0;
    if (nv > 0) {
        for (int i = 0; i + 1 < nv; i++) {
            ext[ne++] = vis[i];
            if (fabs(nc[vis[i + 1]] - nc[vis[i]]) <= a.d_vis)
                for (int j = 1; j < vis[i + 1] - vis[i]; j++) ext[ne++] = vis[i] + j;
        }
        ext[ne++] = vis[nv - 1];
    }
    a.counts[2 * f] = nv; a.counts[2 * f + 1] = ne;
}

__global__ void __launch_bounds__(256) tdlo_vis_compact_kernel(const VisArgs a) {
    __shared__ long long run[2];
    __shared__ long long part[256][2];
    const int tid = threadIdx.x, F = a.n_frames, N = a.n_nodes;
    if (tid == 0) { run[0] = run[1] = 0; a.vis_off[0] = 0; a.ext_off[0] = 0; }
    __syncthreads();
    for (int f0 = 0; f0 < F; f0 += 256) {
        const int f = f0 + tid;
        const long long cv = f < F ? a.counts[2 * f] : 0, ce = f < F ? a.counts[2 * f + 1] : 0;
        part[tid][0] = cv; part[tid][1] = ce;
        __syncthreads();
        if (tid == 0) {                                           // sequential scan of <= 256 entries per round (F is small)
            long long rv = run[0], re = run[1];
            for (int i = 0; i < 256; i++) { const long long v = part[i][0], e = part[i][1]; part[i][0] = rv; part[i][1] = re; rv += v; re += e; }
            run[0] = rv; run[1] = re;
        }
        __syncthreads();
        if (f < F) {
            const long long ov = part[tid][0], oe = part[tid][1];
            a.vis_off[f + 1] = ov + cv; a.ext_off[f + 1] = oe + ce;
            for (int i = 0; i < cv; i++) a.vis[ov + i] = a.tmp_vis[(long long)f * N + i];
            for (int i = 0; i < ce; i++) a.ext[oe + i] = a.tmp_ext[(long long)f * N + i];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// Self-occlusion test (trackdlo_node.cpp:280-343): the edges (i, i+1) are visited nearest-to-camera first; a node whose
// pixel is not yet covered by the thick lines (cv::line, thickness dlo_pixel_width) of the edges visited before is "not
// self-occluded".  No raster here: a node is tested once, at its first visit, against every earlier edge with an exact
// per-pixel restatement of what cv::line draws (OpenCV 4.x drawing.cpp: the segment clipped to the image grown by
// `thickness`, ThickLine = FillConvexPoly of the four corners p +- dp + its Line2 outline + a filled Circle at both ends) --
// oracle/raster.py is the readable version, pinned against cv2.line.  One CTA per frame.
// ------------------------------------------------------------------------------------------
constexpr int SO_SHIFT = 16;
constexpr long long SO_ONE = 1LL << SO_SHIFT, SO_HALF = SO_ONE >> 1;
constexpr int SO_MAX_RADIUS = 256;

struct SoEdge { long long qx[4], qy[4]; int cx[2], cy[2]; int has_quad, drawn; int bx0, bx1, by0, by1; };

__device__ __forceinline__ long long so_trunc(double v) { return (long long)v; }      // (int64)(double): toward zero

// cv::clipLine(Size2l(w, h), pt1, pt2)
__device__ inline bool so_clip_line(long long w, long long h, long long& x1, long long& y1, long long& x2, long long& y2) {
    const long long right = w - 1, bottom = h - 1;
    if (w <= 0 || h <= 0) return false;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        long long a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += so_trunc(__ddiv_rn(__dmul_rn((double)(a - y1), (double)(x2 - x1)), (double)(y2 - y1)));
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += so_trunc(__ddiv_rn(__dmul_rn((double)(a - y2), (double)(x2 - x1)), (double)(y2 - y1)));
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += so_trunc(__ddiv_rn(__dmul_rn((double)(a - x1), (double)(y2 - y1)), (double)(x2 - x1)));
                x1 = a; c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += so_trunc(__ddiv_rn(__dmul_rn((double)(a - x2), (double)(y2 - y1)), (double)(x2 - x1)));
                x2 = a; c2 = 0;
            }
        }
    }
    return (c1 | c2) == 0;
}

// Line2 (16.16 DDA between two fixed-point points, clipped to the image in fixed point): does it set pixel (px, py)?
__device__ inline bool so_line2_hits(int W, int H, long long x1, long long y1, long long x2, long long y2, int px, int py) {
    if (!so_clip_line((long long)W << SO_SHIFT, (long long)H << SO_SHIFT, x1, y1, x2, y2)) return false;
    long long dx = x2 - x1, dy = y2 - y1;
    const long long j = dx < 0 ? -1 : 0, ax = (dx ^ j) - j;
    const long long i = dy < 0 ? -1 : 0, ay = (dy ^ i) - i;
    if (ax > ay) {
        dy = (dy ^ j) - j;
        if (j) { long long t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; }
        if (px == (int)((x2 + SO_HALF) >> SO_SHIFT) && py == (int)((y2 + SO_HALF) >> SO_SHIFT)) return true;     // the far end point is put first
        const long long y_step = (dy * SO_ONE) / (ax | 1);
        const long long ecount = (x2 - x1) >> SO_SHIFT;
        x1 += SO_HALF; y1 += SO_HALF;
        const long long k = (long long)px - (x1 >> SO_SHIFT);
        return k >= 0 && k <= ecount && (int)((y1 + k * y_step) >> SO_SHIFT) == py;
    }
    dx = (dx ^ i) - i;
    if (i) { long long t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; }
    if (px == (int)((x2 + SO_HALF) >> SO_SHIFT) && py == (int)((y2 + SO_HALF) >> SO_SHIFT)) return true;
    const long long x_step = (dx * SO_ONE) / (ay | 1);
    const long long ecount = (y2 - y1) >> SO_SHIFT;
    x1 += SO_HALF; y1 += SO_HALF;
    const long long k = (long long)py - (y1 >> SO_SHIFT);
    return k >= 0 && k <= ecount && (int)((x1 + k * x_step) >> SO_SHIFT) == px;
}

// FillConvexPoly's scan conversion of a quadrilateral: is (px, py) inside the span of row py?
__device__ inline bool so_fill_hits(int W, int H, const long long* vx, const long long* vy, int px, int py) {
    constexpr int npts = 4;
    long long ymin = vy[0], ymax = vy[0], xmin = vx[0], xmax = vx[0];
    int imin = 0;
    for (int i = 0; i < npts; i++) {
        if (vy[i] < ymin) { ymin = vy[i]; imin = i; }
        ymax = vy[i] > ymax ? vy[i] : ymax; xmax = vx[i] > xmax ? vx[i] : xmax; xmin = vx[i] < xmin ? vx[i] : xmin;
    }
    xmin = (xmin + SO_HALF) >> SO_SHIFT; xmax = (xmax + SO_HALF) >> SO_SHIFT;
    ymin = (ymin + SO_HALF) >> SO_SHIFT; ymax = (ymax + SO_HALF) >> SO_SHIFT;
    if (xmax < 0 || ymax < 0 || xmin >= W || ymin >= H) return false;
    if (ymax > H - 1) ymax = H - 1;
    if (py < ymin || py > ymax) return false;
    int e_idx[2] = {imin, imin}, e_di[2] = {1, npts - 1};
    long long e_x[2] = {-SO_ONE, -SO_ONE}, e_dx[2] = {0, 0}, e_ye[2] = {ymin, ymin};
    long long y = ymin;
    int edges = npts;
    for (;;) {
        for (int i = 0; i < 2; i++) {
            if (y >= e_ye[i]) {
                int idx0 = e_idx[i];
                const int di = e_di[i];
                int idx = idx0 + di;
                if (idx >= npts) idx -= npts;
                for (;;) {
                    if (--edges < 0) break;
                    const long long ty = (vy[idx] + SO_HALF) >> SO_SHIFT;
                    if (ty > y) {
                        const long long xs = vx[idx0], xe = vx[idx];
                        e_ye[i] = ty; e_dx[i] = ((xe - xs) * 2 + (ty - y)) / (2 * (ty - y)); e_x[i] = xs; e_idx[i] = idx;
                        break;
                    }
                    idx0 = idx; idx += di;
                    if (idx >= npts) idx -= npts;
                }
            }
        }
        if (edges < 0) return false;
        if (y == py) {
            const int l = e_x[0] <= e_x[1] ? 0 : 1, r = l ^ 1;
            const long long xx1 = (e_x[l] + SO_HALF) >> SO_SHIFT, xx2 = (e_x[r] + SO_HALF) >> SO_SHIFT;
            return px >= xx1 && px <= xx2;          // (xx2 >= 0 && xx1 < W holds for an in-image px inside the span)
        }
        e_x[0] += e_dx[0]; e_x[1] += e_dx[1];
        if (++y > ymax) return false;
    }
}

__global__ void __launch_bounds__(256) tdlo_vis_selfocc_kernel(const VisArgs a) {
    __shared__ SoEdge ed[kMaxNodes];
    __shared__ int pcol[kMaxNodes], prow[kMaxNodes], rank[kMaxNodes], first[kMaxNodes], freef[kMaxNodes];
    __shared__ double dist[kMaxNodes];
    __shared__ int hw[SO_MAX_RADIUS + 1];
    const int f = blockIdx.x, tid = threadIdx.x, N = a.n_nodes, W = a.cols, H = a.rows, th = a.pixel_width;
    const double* Y = a.Y + (long long)f * N * 3;
    const double* P = a.proj + (long long)f * 12;
    const int radius = (int)((((long long)th << (SO_SHIFT - 1)) + SO_HALF) >> SO_SHIFT);
    // filled cv::Circle (midpoint algorithm): half-width of the span at row offset k
    for (int k = tid; k <= radius; k += blockDim.x) hw[k] = -1;
    __syncthreads();
    if (tid == 0) {
        int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
        while (dx >= dy) {
            hw[dy] = max(hw[dy], dx); hw[dx] = max(hw[dx], dy);
            dy++; err += plus; plus += 2;
            const int mask = (err <= 0) - 1;
            err -= minus & mask; dx += mask; minus -= mask & 2;
        }
    }
    // node pixels (:298-313; the Eigen product restated as a sum in index order) and edge midpoints (:283-286)
    for (int m = tid; m < N; m += blockDim.x) {
        const double h[4] = {Y[3 * m], Y[3 * m + 1], Y[3 * m + 2], 1.0};
        double ic[3];
        for (int r = 0; r < 3; r++) {
            double s = 0.0;
            for (int k = 0; k < 4; k++) s = __dadd_rn(s, __dmul_rn(P[4 * r + k], h[k]));
            ic[r] = s;
        }
        const double qx = __ddiv_rn(ic[0], ic[2]), qy = __ddiv_rn(ic[1], ic[2]);
        // static_cast<int>: truncation toward zero; a non-finite or huge quotient (undefined in the reference) lands outside every image
        pcol[m] = (fabs(qx) < 1e9) ? (int)qx : -(1 << 30);
        prow[m] = (fabs(qy) < 1e9) ? (int)qy : -(1 << 30);
        freef[m] = 1; first[m] = 1 << 30;
        if (m + 1 < N) {
            const double mx = __ddiv_rn(__dadd_rn(Y[3 * m], Y[3 * m + 3]), 2.0), my = __ddiv_rn(__dadd_rn(Y[3 * m + 1], Y[3 * m + 4]), 2.0),
                         mz = __ddiv_rn(__dadd_rn(Y[3 * m + 2], Y[3 * m + 5]), 2.0);
            dist[m] = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(mx, mx), __dmul_rn(my, my)), __dmul_rn(mz, mz)));
        }
    }
    __syncthreads();
    // visiting order (std::sort by the midpoint norm; ties, unspecified there, by index) and the edges' geometry
    for (int e = tid; e < N - 1; e += blockDim.x) {
        int rk = 0;
        const double de = dist[e];
        for (int j = 0; j < N - 1; j++) rk += (dist[j] < de) || (dist[j] == de && j < e);
        rank[e] = rk;
        SoEdge& g = ed[e];
        g.drawn = 0; g.has_quad = 0;
        const long long m = th;
        long long x1 = (long long)pcol[e] + m, y1 = (long long)prow[e] + m, x2 = (long long)pcol[e + 1] + m, y2 = (long long)prow[e + 1] + m;
        if (so_clip_line((long long)W + 2 * m, (long long)H + 2 * m, x1, y1, x2, y2)) {       // cv::line: clipped to the image grown by `thickness`
            g.drawn = 1;
            const long long p0x = (x1 - m) << SO_SHIFT, p0y = (y1 - m) << SO_SHIFT, p1x = (x2 - m) << SO_SHIFT, p1y = (y2 - m) << SO_SHIFT;
            const double inv = 1.0 / 65536.0;
            const double dx = __dmul_rn((double)(p0x - p1x), inv), dy = __dmul_rn((double)(p1y - p0y), inv);
            double r = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            const long long thf = (long long)th << (SO_SHIFT - 1);
            long long bx0 = 1LL << 40, bx1 = -(1LL << 40), by0 = 1LL << 40, by1 = -(1LL << 40);
            if (fabs(r) > 2.220446049250313e-16) {
                r = __ddiv_rn(__dadd_rn((double)thf, __dmul_rn((double)((th & 1) * SO_ONE), 0.5)), sqrt(r));
                const long long dpx = __double2ll_rn(__dmul_rn(dy, r)), dpy = __double2ll_rn(__dmul_rn(dx, r));
                g.qx[0] = p0x + dpx; g.qy[0] = p0y + dpy; g.qx[1] = p0x - dpx; g.qy[1] = p0y - dpy;
                g.qx[2] = p1x - dpx; g.qy[2] = p1y - dpy; g.qx[3] = p1x + dpx; g.qy[3] = p1y + dpy;
                g.has_quad = 1;
                for (int i = 0; i < 4; i++) {
                    bx0 = min(bx0, g.qx[i]); bx1 = max(bx1, g.qx[i]); by0 = min(by0, g.qy[i]); by1 = max(by1, g.qy[i]);
                }
            }
            g.cx[0] = (int)((p0x + SO_HALF) >> SO_SHIFT); g.cy[0] = (int)((p0y + SO_HALF) >> SO_SHIFT);
            g.cx[1] = (int)((p1x + SO_HALF) >> SO_SHIFT); g.cy[1] = (int)((p1y + SO_HALF) >> SO_SHIFT);
            // conservative pixel bounding box of everything the line draws (quad corners +- 2 px, circles)
            long long qx0 = (bx0 >> SO_SHIFT) - 2, qx1 = (bx1 >> SO_SHIFT) + 3, qy0 = (by0 >> SO_SHIFT) - 2, qy1 = (by1 >> SO_SHIFT) + 3;
            if (!g.has_quad) { qx0 = qy0 = 1LL << 30; qx1 = qy1 = -(1LL << 30); }
            for (int i = 0; i < 2; i++) {
                qx0 = min(qx0, (long long)g.cx[i] - radius); qx1 = max(qx1, (long long)g.cx[i] + radius);
                qy0 = min(qy0, (long long)g.cy[i] - radius); qy1 = max(qy1, (long long)g.cy[i] + radius);
            }
            g.bx0 = (int)max(qx0, -(1LL << 30)); g.bx1 = (int)min(qx1, 1LL << 30); g.by0 = (int)max(qy0, -(1LL << 30)); g.by1 = (int)min(qy1, 1LL << 30);
        }
    }
    __syncthreads();
    // first visit of every node = the earlier of its two edges
    for (int m = tid; m < N; m += blockDim.x) {
        int fr = 1 << 30;
        if (m > 0) fr = min(fr, rank[m - 1]);
        if (m + 1 < N) fr = min(fr, rank[m]);
        first[m] = fr;
    }
    __syncthreads();
    // (node, earlier edge) pairs
    const int pairs = N * (N - 1);
    for (int idx = tid; idx < pairs; idx += blockDim.x) {
        const int m = idx / (N - 1), e = idx - m * (N - 1);
        if (rank[e] >= first[m]) continue;
        const SoEdge& g = ed[e];
        const int px = pcol[m], py = prow[m];
        if (!g.drawn || px < 0 || px >= W || py < 0 || py >= H) continue;      // pixels outside the image read as 0
        if (px < g.bx0 || px > g.bx1 || py < g.by0 || py > g.by1) continue;
        bool hit = false;
        for (int i = 0; i < 2 && !hit; i++) {
            const int k = abs(py - g.cy[i]);
            hit = k <= radius && hw[k] >= 0 && abs(px - g.cx[i]) <= hw[k];
        }
        if (!hit && g.has_quad) {
            hit = so_fill_hits(W, H, g.qx, g.qy, px, py);
            for (int i = 0; i < 4 && !hit; i++) {
                const int i0 = (i + 3) & 3;
                hit = so_line2_hits(W, H, g.qx[i0], g.qy[i0], g.qx[i], g.qy[i], px, py);
            }
        }
        if (hit) atomicAnd(&freef[m], 0);
    }
    __syncthreads();
    for (int m = tid; m < N; m += blockDim.x) a.free_flag[(long long)f * N + m] = freef[m];
}

// ------------------------------------------------------------------------------------------
// Evaluator frame error (trackdlo/src/evaluator.cpp:233-283 calc_min_distance / get_piecewise_error,
// :333-341 compute_error): mean distance of the nodes of one polyline to the nearest segment of the other, symmetrised.
// One CTA per frame; node distances in parallel, summed in node order by one thread (same order as the reference).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double seg_point_distance(const double* A, const double* B, const double* E) {
    const double ABx = B[0] - A[0], ABy = B[1] - A[1], ABz = B[2] - A[2];
    const double AEx = E[0] - A[0], AEy = E[1] - A[1], AEz = E[2] - A[2];
    const double cx = __dsub_rn(__dmul_rn(AEy, ABz), __dmul_rn(AEz, ABy));
    const double cy = -__dsub_rn(__dmul_rn(AEx, ABz), __dmul_rn(AEz, ABx));
    const double cz = __dsub_rn(__dmul_rn(AEx, ABy), __dmul_rn(AEy, ABx));
    const double abab = __dadd_rn(__dadd_rn(__dmul_rn(ABx, ABx), __dmul_rn(ABy, ABy)), __dmul_rn(ABz, ABz));
    const double aeab = __dadd_rn(__dadd_rn(__dmul_rn(AEx, ABx), __dmul_rn(AEy, ABy)), __dmul_rn(AEz, ABz));
    double distance = __ddiv_rn(sqrt(__dadd_rn(__dadd_rn(__dmul_rn(cx, cx), __dmul_rn(cy, cy)), __dmul_rn(cz, cz))), sqrt(abab));
    const double Px = __dadd_rn(A[0], __ddiv_rn(__dmul_rn(ABx, aeab), abab)), Py = __dadd_rn(A[1], __ddiv_rn(__dmul_rn(ABy, aeab), abab)),
                 Pz = __dadd_rn(A[2], __ddiv_rn(__dmul_rn(ABz, aeab), abab));
    const double APx = Px - A[0], APy = Py - A[1], APz = Pz - A[2];
    const double apab = __dadd_rn(__dadd_rn(__dmul_rn(APx, ABx), __dmul_rn(APy, ABy)), __dmul_rn(APz, ABz));
    if (apab < 0 || apab > abab) {
        const double BEx = E[0] - B[0], BEy = E[1] - B[1], BEz = E[2] - B[2];
        const double dAE = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(AEx, AEx), __dmul_rn(AEy, AEy)), __dmul_rn(AEz, AEz)));
        const double dBE = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(BEx, BEx), __dmul_rn(BEy, BEy)), __dmul_rn(BEz, BEz)));
        distance = dAE > dBE ? dBE : dAE;
    }
    return distance;
}

__global__ void __launch_bounds__(256) tdlo_tracking_error_kernel(int n_track, int n_true, const double* Yt_all, const double* Yr_all, double* err) {
    __shared__ double yt[kMaxNodes * 3], yr[kMaxNodes * 3], d1[kMaxNodes], d2[kMaxNodes];
    const int f = blockIdx.x, tid = threadIdx.x;
    const double* Yt = Yt_all + (long long)f * n_track * 3;
    const double* Yr = Yr_all + (long long)f * n_true * 3;
    for (int i = tid; i < 3 * n_track; i += blockDim.x) yt[i] = Yt[i];
    for (int i = tid; i < 3 * n_true; i += blockDim.x) yr[i] = Yr[i];
    __syncthreads();
    for (int idx = tid; idx < n_track; idx += blockDim.x) {
        double dist = -1;
        for (int i = 0; i < n_true - 1; i++) { const double di = seg_point_distance(yr + 3 * i, yr + 3 * (i + 1), yt + 3 * idx); if (dist == -1 || di < dist) dist = di; }
        d1[idx] = dist;
    }
    for (int idx = tid; idx < n_true; idx += blockDim.x) {
        double dist = -1;
        for (int i = 0; i < n_track - 1; i++) { const double di = seg_point_distance(yt + 3 * i, yt + 3 * (i + 1), yr + 3 * idx); if (dist == -1 || di < dist) dist = di; }
        d2[idx] = dist;
    }
    __syncthreads();
    if (tid == 0) {
        double t1 = 0.0, t2 = 0.0;
        for (int i = 0; i < n_track; i++) t1 += d1[i];
        for (int i = 0; i < n_true; i++) t2 += d2[i];
        err[f] = (t1 / n_track + t2 / n_true) / 2;
    }
}

}  // namespace tdlo
