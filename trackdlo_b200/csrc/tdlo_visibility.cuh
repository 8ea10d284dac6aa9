// Visibility front-end feeding tracking_step, batched over independent frames:
//   shortest node-to-point distances           trackdlo/src/trackdlo_node.cpp:254-277
//   visible_nodes (sorted) / visible_nodes_extended (d_vis rule)   trackdlo_node.cpp:346-360
// The self-occlusion raster (:280-343) is out of scope: every node counts as not self-occluded.
// Four small kernels, no host round trip (the device-pointer entry point is fully stream-ordered):
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
        if (d <= a.tau) vis[nv++] = m;
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
