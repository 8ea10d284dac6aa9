// Perception front-end that produces the tracker's input cloud, batched over independent frames:
//   BGR -> HSV (8-bit OpenCV formula) -> inRange band(s) -> AND with the grey occlusion mask        trackdlo_node.cpp:159-180, 88-119
//   mask + depth (uint16 mm) -> float32 points through the projection matrix                         trackdlo_node.cpp:195-233
//   pcl::VoxelGrid centroid down-sampling (leaf 0.008), float32 -> double                             trackdlo_node.cpp:236-242
// The output (X concatenated over the frames + CSR offsets, in device memory) is exactly the input of
// tdlo_visibility_batched_device / tdlo_tracking_step_batched_device: image in -> nodes out without a host hop.
//
// Integer / byte work, HBM-bound (5 B per pixel read twice); no sort: a frame's voxels live in a dense grid over the
// bounding box of its masked points (what PCL's linear voxel index addresses), centroids are accumulated with 64-bit
// INTEGER atomics (fixed point 2^-36 m: exact for float32 inputs, order-independent => bit-deterministic), and the
// occupied cells are emitted in ascending cell order (PCL's output order) by a tiled scan.
// Semantics and where they depart from PCL's float32 accumulation order: oracle/frontend.py (the checker).
#pragma once

#include "tdlo_common.cuh"

namespace tdlo {

constexpr int FE_TILE = 1024;                 // grid cells per tile of the emit scan
constexpr double FE_FIX = 68719476736.0;      // 2^36
constexpr int FE_ST_EMPTY = 1, FE_ST_GRID = 2, FE_ST_CAPACITY = 4;

struct FeArgs {
    int n_frames, rows, cols, multi;
    const unsigned char* bgr; const unsigned short* depth; const unsigned char* occl;   // occl may be null
    const double* proj;                       // [F][12] row-major 3x4 projection matrix (only fx, fy, cx, cy are read)
    int lo[3], hi[3];
    float inv_leaf;
    int* bbox;                                // [F][6] min ijk, max ijk
    int* dims;                                // [F][4] dx, dy, dz, status
    long long* cell_base;                     // [F+1]
    long long* tile_base;                     // [F+1]
    long long* acc;                           // [cells_cap][4] sum x, y, z (fixed point), count
    long long cells_cap;
    int* tile_cnt; long long* tile_off; long long tiles_cap;
    double* X; long long* x_off; long long x_cap; int* status;
};

// cv::cvtColor(COLOR_BGR2HSV) for 8-bit pixels (OpenCV RGB2HSV_b: hsv_shift = 12, hrange = 180)
__device__ __forceinline__ void fe_bgr2hsv(int b, int g, int r, int& h, int& s, int& v) {
    v = max(max(b, g), r);
    const int vmin = min(min(b, g), r), diff = v - vmin;
    // sdiv_table[v] = round((255 << 12) / v), hdiv_table[diff] = round((180 << 12) / (6 diff)); exact in double
    const int sdiv = v ? __double2int_rn(1044480.0 / (double)v) : 0;
    const int hdiv = diff ? __double2int_rn(737280.0 / (6.0 * (double)diff)) : 0;
    s = (diff * sdiv + (1 << 11)) >> 12;
    int hh = (v == r) ? (g - b) : ((v == g) ? (b - r + 2 * diff) : (r - g + 4 * diff));
    hh = (hh * hdiv + (1 << 11)) >> 12;
    h = hh + (hh < 0 ? 180 : 0);
}
__device__ __forceinline__ bool fe_in(int h, int s, int v, int l0, int l1, int l2, int u0, int u1, int u2) {
    return h >= l0 && h <= u0 && s >= l1 && s <= u1 && v >= l2 && v <= u2;
}

// mask of one pixel and, if set, its float32 point and integer voxel coordinates
__device__ __forceinline__ bool fe_pixel(const FeArgs& a, int f, long long pix, float& x, float& y, float& z, int& ix, int& iy, int& iz) {
    const long long np = (long long)a.rows * a.cols, g = (long long)f * np + pix;
    const unsigned char* c = a.bgr + g * 3;
    int h, s, v;
    fe_bgr2hsv(c[0], c[1], c[2], h, s, v);
    bool m;
    if (a.multi) {          // color_thresholding (trackdlo_node.cpp:88-119): blue | red_1 | red_2 | yellow
        m = fe_in(h, s, v, 90, 90, 60, 130, 255, 255) || fe_in(h, s, v, 130, 60, 50, 255, 255, 255) ||
            fe_in(h, s, v, 0, 60, 50, 10, 255, 255) || fe_in(h, s, v, 15, 100, 80, 40, 255, 255);
    } else m = fe_in(h, s, v, a.lo[0], a.lo[1], a.lo[2], a.hi[0], a.hi[1], a.hi[2]);
    if (m && a.occl) {      // mask & grey(occlusion image): cv::COLOR_BGR2GRAY 8-bit, then bitwise_and with 255
        const unsigned char* o = a.occl + g * 3;
        const int grey = (o[0] * 3735 + o[1] * 19235 + o[2] * 9798 + (1 << 14)) >> 15;
        m = grey != 0;
    }
    if (!m) return false;
    const double* P = a.proj + (long long)f * 12;
    const double fx = P[0], cx = P[2], fy = P[5], cy = P[6];
    const int i = (int)(pix / a.cols), j = (int)(pix - (long long)i * a.cols);
    const double pz = __ddiv_rn((double)a.depth[g], 1000.0);
    x = (float)__ddiv_rn(__dmul_rn((double)j - cx, pz), fx);          // trackdlo_node.cpp:218-220, stored into float fields
    y = (float)__ddiv_rn(__dmul_rn((double)i - cy, pz), fy);
    z = (float)pz;
    ix = (int)floorf(__fmul_rn(x, a.inv_leaf)); iy = (int)floorf(__fmul_rn(y, a.inv_leaf)); iz = (int)floorf(__fmul_rn(z, a.inv_leaf));
    return true;
}

__global__ void __launch_bounds__(256) fe_bbox_kernel(const FeArgs a) {
    const int f = blockIdx.y;
    const long long np = (long long)a.rows * a.cols;
    int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
    for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < np; pix += (long long)gridDim.x * blockDim.x) {
        float x, y, z; int ix, iy, iz;
        if (fe_pixel(a, f, pix, x, y, z, ix, iy, iz)) {
            mn[0] = min(mn[0], ix); mn[1] = min(mn[1], iy); mn[2] = min(mn[2], iz);
            mx[0] = max(mx[0], ix); mx[1] = max(mx[1], iy); mx[2] = max(mx[2], iz);
        }
    }
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const int lo = __reduce_min_sync(0xffffffffu, mn[d]), hi = __reduce_max_sync(0xffffffffu, mx[d]);
        if ((threadIdx.x & 31) == 0) {
            if (lo != INT_MAX) atomicMin(a.bbox + f * 6 + d, lo);
            if (hi != INT_MIN) atomicMax(a.bbox + f * 6 + 3 + d, hi);
        }
    }
}

// one CTA: grid dimensions, cell / tile bases of every frame (frames are few: sequential scan by thread 0)
__global__ void fe_layout_kernel(const FeArgs a) {
    if (threadIdx.x != 0) return;
    long long cells = 0, tiles = 0;
    for (int f = 0; f < a.n_frames; f++) {
        const int* b = a.bbox + f * 6;
        int st = 0;
        long long dx = 0, dy = 0, dz = 0, n = 0;
        if (b[0] == INT_MAX) st = FE_ST_EMPTY;
        else {
            dx = (long long)b[3] - b[0] + 1; dy = (long long)b[4] - b[1] + 1; dz = (long long)b[5] - b[2] + 1;
            n = dx * dy * dz;
            // PCL refuses grids whose index overflows an int (and returns the cloud unfiltered); here the frame is refused
            if (n > (long long)INT_MAX || cells + n > a.cells_cap || tiles + (n + FE_TILE - 1) / FE_TILE > a.tiles_cap) { st = FE_ST_GRID; n = 0; }
        }
        a.dims[f * 4] = (int)dx; a.dims[f * 4 + 1] = (int)dy; a.dims[f * 4 + 2] = (int)dz; a.dims[f * 4 + 3] = st;
        a.cell_base[f] = cells; a.tile_base[f] = tiles;
        cells += n; tiles += (n + FE_TILE - 1) / FE_TILE;
    }
    a.cell_base[a.n_frames] = cells; a.tile_base[a.n_frames] = tiles;
}

__global__ void __launch_bounds__(256) fe_clear_kernel(const FeArgs a) {
    const long long n4 = a.cell_base[a.n_frames] * 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) a.acc[i] = 0;
}

__global__ void __launch_bounds__(256) fe_accum_kernel(const FeArgs a) {
    const int f = blockIdx.y;
    if (a.dims[f * 4 + 3]) return;
    const long long np = (long long)a.rows * a.cols;
    const int* b = a.bbox + f * 6;
    const long long dx = a.dims[f * 4], dy = a.dims[f * 4 + 1], base = a.cell_base[f];
    for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < np; pix += (long long)gridDim.x * blockDim.x) {
        float x, y, z; int ix, iy, iz;
        if (fe_pixel(a, f, pix, x, y, z, ix, iy, iz)) {
            const long long cell = base + (ix - b[0]) + (iy - b[1]) * dx + (iz - b[2]) * dx * dy;
            unsigned long long* c = reinterpret_cast<unsigned long long*>(a.acc + cell * 4);
            atomicAdd(c, (unsigned long long)__double2ll_rn((double)x * FE_FIX));
            atomicAdd(c + 1, (unsigned long long)__double2ll_rn((double)y * FE_FIX));
            atomicAdd(c + 2, (unsigned long long)__double2ll_rn((double)z * FE_FIX));
            atomicAdd(c + 3, 1ull);
        }
    }
}

// occupied cells per tile
__global__ void __launch_bounds__(256) fe_count_kernel(const FeArgs a) {
    __shared__ int red[8];
    const long long total_tiles = a.tile_base[a.n_frames];
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int f = 0;
        while (a.tile_base[f + 1] <= tile) f++;                      // frames are few
        const long long n = a.cell_base[f + 1] - a.cell_base[f];
        const long long c0 = (tile - a.tile_base[f]) * FE_TILE;
        int cnt = 0;
        for (int i = threadIdx.x; i < FE_TILE; i += blockDim.x) {
            const long long c = c0 + i;
            if (c < n && a.acc[(a.cell_base[f] + c) * 4 + 3] != 0) cnt++;
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w]; a.tile_cnt[tile] = t; }
        __syncthreads();
    }
}

// one CTA: exclusive scan of the tile counts -> point offsets; per-frame offsets, capacity check, status
__global__ void __launch_bounds__(256) fe_scan_kernel(const FeArgs a) {
    __shared__ long long part[256];
    __shared__ long long run;
    const int tid = threadIdx.x;
    const long long total_tiles = a.tile_base[a.n_frames];
    if (tid == 0) run = 0;
    __syncthreads();
    for (long long t0 = 0; t0 < total_tiles; t0 += 256) {
        const long long t = t0 + tid;
        const long long c = t < total_tiles ? a.tile_cnt[t] : 0;
        part[tid] = c;
        __syncthreads();
        if (tid == 0) { long long r = run; for (int i = 0; i < 256; i++) { const long long v = part[i]; part[i] = r; r += v; } run = r; }
        __syncthreads();
        if (t < total_tiles) a.tile_off[t] = part[tid];
        __syncthreads();
    }
    if (tid == 0) {
        const long long total = run;
        long long e = 0;
        bool full = false;
        for (int f = 0; f < a.n_frames; f++) {      // frames are emitted in order; the first that does not fit and all later ones get no points
            const long long o0 = a.tile_base[f] < total_tiles ? a.tile_off[a.tile_base[f]] : total;
            const long long o1 = a.tile_base[f + 1] < total_tiles ? a.tile_off[a.tile_base[f + 1]] : total;
            int st = a.dims[f * 4 + 3];
            if (!full && o1 <= a.x_cap) { a.x_off[f] = o0; e = o1; }
            else { full = true; a.x_off[f] = e; if (o1 > o0) st |= FE_ST_CAPACITY; }
            if (a.status) a.status[f] = st;
            a.dims[f * 4 + 3] = st;
        }
        a.x_off[a.n_frames] = e;
    }
}

__global__ void fe_init_kernel(const FeArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n_frames * 6) a.bbox[i] = (i % 6) < 3 ? INT_MAX : INT_MIN;
}

// centroids of the occupied cells of a tile, in cell order
__global__ void __launch_bounds__(256) fe_emit_kernel(const FeArgs a) {
    __shared__ int wsum[8];
    __shared__ int tile_run;
    const long long total_tiles = a.tile_base[a.n_frames];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int f = 0;
        while (a.tile_base[f + 1] <= tile) f++;
        if (a.dims[f * 4 + 3] & FE_ST_CAPACITY) continue;           // uniform over the CTA
        const long long n = a.cell_base[f + 1] - a.cell_base[f], cb = a.cell_base[f];
        const long long c0 = (tile - a.tile_base[f]) * FE_TILE;
        const long long out0 = a.tile_off[tile];
        if (tid == 0) tile_run = 0;
        __syncthreads();
        for (int i0 = 0; i0 < FE_TILE; i0 += 256) {                 // 4 rounds of 256 consecutive cells
            const long long c = c0 + i0 + tid;
            const long long* cell = a.acc + (cb + c) * 4;
            const long long cnt = c < n ? cell[3] : 0;
            const unsigned bal = __ballot_sync(0xffffffffu, cnt != 0);
            if (lane == 0) wsum[warp] = __popc(bal);
            __syncthreads();
            int before = tile_run;
            for (int w = 0; w < warp; w++) before += wsum[w];
            if (cnt != 0) {
                const long long o = out0 + before + __popc(bal & ((1u << lane) - 1u));
                const double inv = 1.0 / FE_FIX, dn = (double)cnt;
                // centroid = exact mean of the float32 coordinates, rounded once to float32 (PCL's point type), widened (:242)
                a.X[o * 3] = (double)(float)(__ddiv_rn((double)cell[0], dn) * inv);
                a.X[o * 3 + 1] = (double)(float)(__ddiv_rn((double)cell[1], dn) * inv);
                a.X[o * 3 + 2] = (double)(float)(__ddiv_rn((double)cell[2], dn) * inv);
            }
            __syncthreads();
            if (tid == 0) { int t = 0; for (int w = 0; w < 8; w++) t += wsum[w]; tile_run += t; }
            __syncthreads();
        }
    }
}

}  // namespace tdlo
