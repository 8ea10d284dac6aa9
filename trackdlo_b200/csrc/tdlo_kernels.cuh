// Device side of the B200-native TrackDLO registration path (sm_100a).
//
// One thread-block CLUSTER owns one frame for its whole life: prune -> set-up -> all EM
// iterations of cpd_lle (trackdlo/src/trackdlo.cpp:161-441), and in tracking mode the full
// tracking_step (trackdlo.cpp:900-999: pre-processing registration, traverse_euclidean, main
// registration) without returning to the host.  Clusters pull frames from an atomic queue
// (persistent scheduling), so data-dependent iteration counts (trackdlo.cpp:424-428) balance
// automatically.  The Nn x Mp affinity matrix P is never written to HBM: each CTA streams its
// slice of the frame's points in tiles, keeps one P tile in shared memory and reduces it to
// P1 / PX partial sums in registers.
//
// Everything is fp64 (the reference is MatrixXd end to end; the parity gate is 1e-5 on W).
#pragma once

#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include <math.h>

namespace cg = cooperative_groups;

namespace tdlo {

constexpr int kMaxThreads = 256;
constexpr int kMaxCluster = 16;
constexpr int kMaxNodes = 256;

// status bits (mirror include/trackdlo_b200.h)
constexpr int ST_NOT_CONVERGED = 1, ST_SINGULAR = 2, ST_TOO_FEW_NODES = 4, ST_EMPTY = 8, ST_TRAVERSE_UB = 16,
              ST_PRE_NOT_CONVERGED = 32;

struct CpdP {
    double beta, lambda, gamma, mu, tol, alpha, k_vis, tau, prune_radius;
    int max_iter, include_lle;
};

// ------------------------------------------------------------------------------------------
// shared memory layout (bytes)
// ------------------------------------------------------------------------------------------
struct SmemL {
    int tab, node4, wbuf, y0, s, vw, yext, jd, hy0, p1, px, wsol, tnew, pacc, red, gjbuf, prow, used, ptile, total;
};
__host__ __device__ inline SmemL smem_layout(int N, int tile) {
    SmemL l;
    int o = 0;
    l.tab = o; o += 64 * 8;
    l.node4 = o; o += N * 32;
    l.wbuf = o; o += tile * 32;
    l.y0 = o; o += 3 * N * 8;
    l.s = o; o += N * 8;
    l.vw = o; o += N * 8;
    l.yext = o; o += 3 * N * 8;
    l.jd = o; o += N * 8;
    l.hy0 = o; o += 3 * N * 8;
    l.p1 = o; o += N * 8;
    l.px = o; o += 3 * N * 8;
    l.wsol = o; o += 3 * N * 8;
    l.tnew = o; o += 3 * N * 8;
    l.pacc = o; o += 4 * N * 8;
    l.red = o; o += 64 * 8;
    l.gjbuf = o; o += 132 * 8;
    l.prow = o; o += N * 4;
    l.used = o; o += N * 4;
    o = (o + 31) & ~31;
    l.ptile = o; o += (tile / 32) * N * 33 * 8;      // one [N][33] P slice per warp
    l.total = o;
    return l;
}

struct KArgs {
    int mode;            // 0 = batched cpd_lle, 1 = batched tracking_step
    int n_frames;
    int node_stride;     // row stride of Y / priors / W / H (nodes)
    int tile;            // points per tile == blockDim.x
    int nmax;            // largest node count in the batch (sizes shared memory)
    // frame data (device pointers)
    const double* X; const long long* x_off;
    const int* n_nodes;
    double* Y; double* sigma2;
    const double* priors; const int* n_priors; const int* n_visible;
    const double* H;
    double* W; int* iters; int* status;
    // tracking-step extras
    const double* rest;
    const int* vis; const long long* vis_off;
    const int* ext; const long long* ext_off;
    double* guide_out; double* priors_out; int* n_priors_out; int* state_out;
    CpdP p0;             // mode 0: the call's params; mode 1: pre-processing registration
    CpdP p1;             // mode 1: main registration
    // workspace
    double* Xc;          // compacted points, same indexing as X
    unsigned short* bkt; // nearest-node bucket per raw point (sort key), same indexing as X
    double* scratch;     // per-cluster scratch
    long long scratch_stride;   // doubles per cluster
    int* queue;          // frame queue counter
    int scr_nodes;       // node capacity the scratch layout was sized for
    unsigned long long* prof;   // optional [16] phase cycle counters (rank 0: 0..7, other ranks: 8..15)
    SmemL L;             // shared-memory layout, computed by the host (keeps address arithmetic out of the kernel)
};

__constant__ double c_exp_tab[64];   // 2^(j/64), filled by the host

// ------------------------------------------------------------------------------------------
// per-cluster global scratch layout (doubles); N = scr_nodes
// ------------------------------------------------------------------------------------------
struct Scr {
    long long G, HG, H, AB, PART, DMIN, GATH, STATE, TRV, PRI, GUIDE, CTL, total;
};
__host__ __device__ inline Scr scr_layout(int N) {
    Scr s;
    long long o = 0, n2 = (long long)N * N;
    s.G = o; o += n2;
    s.HG = o; o += n2;
    s.H = o; o += n2;
    s.AB = o; o += (long long)N * (N + 4);
    s.PART = o; o += (long long)kMaxCluster * (4 * N + 4);
    s.DMIN = o; o += (long long)kMaxCluster * N;
    s.GATH = o; o += kMaxCluster * 2;
    s.STATE = o; o += 3 * N + 8;
    s.TRV = o; o += 2LL * (N + 2) * 4;
    s.PRI = o; o += (2LL * N + 4) * 4;
    s.GUIDE = o; o += 3 * N;
    s.CTL = o; o += 8;
    s.total = (o + 15) & ~15LL;
    return s;
}

struct Smem {
    double* tab; double4* node4; double4* wbuf;
    double *y0, *s, *vw, *yext, *jd, *hy0, *p1, *px, *wsol, *tnew, *pacc, *red, *gjbuf;
    int *prow, *used;
    double* ptile;
};

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the block; every thread gets the total.  `red` needs >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < nw; w++) t += red[w];
    return t;
}

// exp(-z) for z >= 0, fp64: 64-entry table 2^(j/64) + degree-5 polynomial, relative error
// <= 3e-16 + |z| * 5e-17 (the ln2/64 reduction constant is a single double).
// z is clamped to [0, ~706]: entries that the reference would compute as < 2.4e-307 (down to
// denormals / exact 0) come out as ~2.4e-307 here -- 290 orders of magnitude below the outlier
// constant c they are added to, i.e. no effect on any result bit that survives the division.
__device__ __forceinline__ double exp_neg(double z, const double* __restrict__ tab) {
    const double L = 92.33248261689366;           // 64 / ln 2
    const double C_HI = 0.010830424696249145;     // ln 2 / 64
    const double MAGIC = 6755399441055744.0;      // 1.5 * 2^52
    z = __hiloint2double(min(__double2hiint(z), 0x40861000), __double2loint(z));   // NaN also lands here
    const double t = fma(z, -L, MAGIC);
    const int n = __double2loint(t);
    const double nf = t - MAGIC;
    const double r = fma(nf, -C_HI, -z);
    const double tj = tab[n & 63];
    const double r2 = r * r;
    double q = fma(r, 8.3333333333333332e-3, 4.1666666666666664e-2);
    q = fma(q, r, 1.6666666666666666e-1);
    q = fma(q, r, 0.5);
    const double p = fma(q, r2, r);
    const double e = fma(tj, p, tj);
    return __hiloint2double(__double2hiint(e) + ((n >> 6) << 20), __double2loint(e));
}

__device__ __forceinline__ double dist2(double ax, double ay, double az, double bx, double by, double bz) {
    const double dx = ax - bx, dy = ay - by, dz = az - bz;
    return dx * dx + dy * dy + dz * dz;
}

// ------------------------------------------------------------------------------------------
// Prune this CTA's slice of the frame (trackdlo.cpp:177-195), accumulate the sum of squared
// distances of the kept points to all nodes (sigma2 init, :263-273), and write the kept points
// to Xc[r0 ...) STABLY SORTED BY NEAREST NODE (counting sort, deterministic).  Sorting changes only
// the summation order of the E-step reductions; it makes the points of a warp neighbours along
// the DLO, which is what lets the E-step skip node ranges whose P entries are exactly 0.
// bkt: global temp, one uint16 per raw point (nearest node, 0xffff = pruned).
// Returns the number of kept points (uniform over the block); *sum_out gets the block sum.
// ------------------------------------------------------------------------------------------
__device__ int prune_sort_slice(const Smem& sm, const double* __restrict__ Xraw, long long r0, long long r1,
                                double* __restrict__ Xc, unsigned short* __restrict__ bkt, int Nn, double radius,
                                double* sum_out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5, nt = blockDim.x;
    int* hist = reinterpret_cast<int*>(sm.ptile);      // [Nn+1] counts, then running bucket offsets
    int* cnt = hist + 264;                             // [nw][Nn] per-warp counts of the current tile
    for (int i = tid; i <= Nn; i += nt) hist[i] = 0;
    for (int i = tid; i < nw * Nn; i += nt) cnt[i] = 0;
    __syncthreads();
    double sum2 = 0.0;
    for (long long base = r0; base < r1; base += nt) {
        const long long n = base + tid;
        const bool valid = n < r1;
        double x = 0, y = 0, z = 0;
        if (valid) { x = __ldg(Xraw + n * 3); y = __ldg(Xraw + n * 3 + 1); z = __ldg(Xraw + n * 3 + 2); }
        double best = 1e300, tot = 0.0;
        int a = 0;
        for (int j = 0; j < Nn; j++) {
            const double4 q = sm.node4[j];
            const double d2 = dist2(q.x, q.y, q.z, x, y, z);
            tot += d2;
            if (d2 < best) { best = d2; a = j; }
        }
        const bool keep = valid && (sqrt(best) < radius);
        if (valid) bkt[n] = keep ? (unsigned short)a : (unsigned short)0xffff;
        if (keep) { atomicAdd(&hist[a], 1); sum2 += tot; }
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int bk = 0; bk < Nn; bk++) { const int c = hist[bk]; hist[bk] = run; run += c; }
        hist[Nn] = run;
    }
    __syncthreads();
    const int count = hist[Nn];
    for (long long base = r0; base < r1; base += nt) {
        const long long n = base + tid;
        const bool valid = n < r1;
        const unsigned bk = valid ? (unsigned)bkt[n] : 0xffffu;
        const bool keep = bk != 0xffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, bk);
        const int rk = __popc(peers & ((1u << lane) - 1u));
        if (keep && rk == 0) cnt[warp * Nn + bk] = __popc(peers);
        __syncthreads();
        if (keep) {
            int off = hist[bk] + rk;
            for (int w = 0; w < warp; w++) off += cnt[w * Nn + bk];
            const long long dst = r0 + off;
            Xc[dst * 3] = __ldg(Xraw + n * 3); Xc[dst * 3 + 1] = __ldg(Xraw + n * 3 + 1); Xc[dst * 3 + 2] = __ldg(Xraw + n * 3 + 2);
        }
        __syncthreads();
        for (int bb = tid; bb < Nn; bb += nt) {
            int sc = 0;
            for (int w = 0; w < nw; w++) { sc += cnt[w * Nn + bb]; cnt[w * Nn + bb] = 0; }
            hist[bb] += sc;
        }
        __syncthreads();
    }
    *sum_out = block_sum(sum2, sm.red);
    return count;
}

// ------------------------------------------------------------------------------------------
// Visibility pre-pass: per-node min squared distance to this CTA's points (trackdlo.cpp:279-296).
// Result (per-CTA partial) is written to dmin_out[0..Nn).
// ------------------------------------------------------------------------------------------
__device__ void dmin_slice(const Smem& sm, const double* __restrict__ Xc, int n_local, int Nn, double* dmin_out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5, tile = blockDim.x;
    double* wmin = sm.ptile;                   // [nw][Nn], P tile is idle here
    for (int i = tid; i < nw * Nn; i += tile) wmin[i] = 1e300;
    __syncthreads();
    for (int base = 0; base < n_local; base += tile) {
        const int n = base + tid;
        const bool valid = n < n_local;
        double x = 0, y = 0, z = 0;
        if (valid) { x = Xc[(long long)n * 3]; y = Xc[(long long)n * 3 + 1]; z = Xc[(long long)n * 3 + 2]; }
        for (int j = 0; j < Nn; j++) {
            const double4 q = sm.node4[j];
            double d2 = dist2(q.x, q.y, q.z, x, y, z);
            if (!valid) d2 = 1e300;
            // warp min of a non-negative double: order == order of its bit pattern
            const unsigned hi = (unsigned)__double2hiint(d2);
            const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
            const unsigned lo = (hi == mh) ? (unsigned)__double2loint(d2) : 0xffffffffu;
            const unsigned ml = __reduce_min_sync(0xffffffffu, lo);
            if (lane == 0) {
                const double m = __hiloint2double((int)mh, (int)ml);
                double* slot = wmin + warp * Nn + j;
                if (m < *slot) *slot = m;
            }
        }
    }
    __syncthreads();
    for (int j = tid; j < Nn; j += tile) {
        double m = wmin[j];
        for (int w = 1; w < nw; w++) m = fmin(m, wmin[w * Nn + j]);
        __stcg(dmin_out + j, m);
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// Fused E-step over this CTA's slice (trackdlo.cpp:278-389): distances -> arg-max node ->
// geodesic distances -> P -> (visibility weights) -> normalisation -> P1, PX, sum Pt1*|x|^2.
// Every WARP is autonomous: it takes 32 points at a time (lane = point), writes their P columns
// into its private shared-memory slice pt[node][33] (phase A), then switches to lane = node and
// accumulates P1/PX for its nodes over those 32 points in registers (phase B).  Only __syncwarp
// separates the phases, so the warps of a CTA overlap their phases freely and no block barrier
// sits in the hot loop.  NPASS = ceil(Nn / 32) node passes in phase B (compile time).
// part_out: [Nn][4] = {P1, PX.x, PX.y, PX.z}, then [4*Nn] = sum_n Pt1_n |x_n|^2.
// ------------------------------------------------------------------------------------------
template <int NPASS, bool VIS>
__device__ void estep_slice(const Smem& sm, const double* __restrict__ Xc, int n_local, int Nn,
                            double sigma2, double c_norm, double rscale, double* part_out) {
    constexpr int RS = 33;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5, nt = blockDim.x;
    double* __restrict__ pt = sm.ptile + warp * (Nn * RS);
    double4* __restrict__ wb = sm.wbuf + warp * 32;
    double* __restrict__ pcol = pt + lane;
    const double* __restrict__ tab = sm.tab;
    double acc[NPASS][4];
#pragma unroll
    for (int ps = 0; ps < NPASS; ps++) { acc[ps][0] = acc[ps][1] = acc[ps][2] = acc[ps][3] = 0.0; }
    double sxx = 0.0;
    const double uflow = 1490.2 * sigma2;      // cheap pre-test: below this exp(-0.5*d2/sigma2) cannot underflow to 0

    for (int base = warp * 32; base < n_local; base += nw * 32) {
        const int n = base + lane;
        const bool valid = n < n_local;
        double x = 0, y = 0, z = 0;
        if (valid) { x = Xc[(long long)n * 3]; y = Xc[(long long)n * 3 + 1]; z = Xc[(long long)n * 3 + 2]; }

        // ---- nearest node (== arg-max of the Gaussian P of trackdlo.cpp:298-310).
        // Exact pruning of the search range: with c = the point of lane 0 and rho = max_lane |x - c|, a node m
        // with |Y_m - c| > min_m' |Y_m' - c| + 2 rho is strictly farther from EVERY point of this warp than the
        // node nearest to c (triangle inequality), so it can be neither the arg-min nor tie with it.  The scan
        // covers the contiguous index range [ja, jb] spanned by the surviving nodes -> same first-minimum as
        // the full scan.  The points are sorted by nearest node, so the range is a handful of nodes.
        int ja, jb;
        {
            const double cx = __shfl_sync(0xffffffffu, x, 0), cy = __shfl_sync(0xffffffffu, y, 0), cz = __shfl_sync(0xffffffffu, z, 0);
            const double r2 = valid ? dist2(x, y, z, cx, cy, cz) : 0.0;
            const unsigned rh = __reduce_max_sync(0xffffffffu, (unsigned)__double2hiint(r2));
            const unsigned rl = __reduce_max_sync(0xffffffffu, (unsigned)__double2hiint(r2) == rh ? (unsigned)__double2loint(r2) : 0u);
            const double rho = sqrt(__hiloint2double((int)rh, (int)rl));
            double dc[NPASS];
            double dloc = 1e300;
#pragma unroll
            for (int ps = 0; ps < NPASS; ps++) {
                const int m = lane + 32 * ps;
                dc[ps] = 1e300;
                if (m < Nn) { const double4 q = sm.node4[m]; dc[ps] = sqrt(dist2(q.x, q.y, q.z, cx, cy, cz)); }
                dloc = fmin(dloc, dc[ps]);
            }
            const unsigned dh = __reduce_min_sync(0xffffffffu, (unsigned)__double2hiint(dloc));
            const unsigned dl = __reduce_min_sync(0xffffffffu, (unsigned)__double2hiint(dloc) == dh ? (unsigned)__double2loint(dloc) : 0xffffffffu);
            const double lim = (__hiloint2double((int)dh, (int)dl) + 2.0 * rho) * (1.0 + 1e-12) + 1e-300;
            ja = Nn; jb = -1;
#pragma unroll
            for (int ps = 0; ps < NPASS; ps++) {
                const unsigned mk = __ballot_sync(0xffffffffu, dc[ps] <= lim);
                if (mk) { if (ja == Nn) ja = 32 * ps + __ffs(mk) - 1; jb = 32 * ps + 31 - __clz(mk); }
            }
        }
        double best = 1e300;
        int a = ja;
        {
            int j = ja;
            for (; j + 3 <= jb; j += 4) {
                const double4 q0 = sm.node4[j], q1 = sm.node4[j + 1], q2 = sm.node4[j + 2], q3 = sm.node4[j + 3];
                const double e0 = dist2(q0.x, q0.y, q0.z, x, y, z), e1 = dist2(q1.x, q1.y, q1.z, x, y, z);
                const double e2 = dist2(q2.x, q2.y, q2.z, x, y, z), e3 = dist2(q3.x, q3.y, q3.z, x, y, z);
                if (e0 < best) { best = e0; a = j; }
                if (e1 < best) { best = e1; a = j + 1; }
                if (e2 < best) { best = e2; a = j + 2; }
                if (e3 < best) { best = e3; a = j + 3; }
            }
            for (; j <= jb; j++) {
                const double4 q = sm.node4[j];
                const double d2 = dist2(q.x, q.y, q.z, x, y, z);
                if (d2 < best) { best = d2; a = j; }
            }
        }
        // whole column underflows to 0 in the reference -> maxCoeff returns index 0
        if (best > uflow && (-0.5 * best) / sigma2 < -745.1332191019412) a = 0;
        int q1 = a - 1; if (q1 == -1) q1 = 2;
        int q2 = a + 1; if (q2 == Nn) q2 = Nn - 3;
        double da, d1, d2n;
        { const double4 q = sm.node4[a];  da  = sqrt(dist2(q.x, q.y, q.z, x, y, z)); }
        { const double4 q = sm.node4[q1]; d1  = sqrt(dist2(q.x, q.y, q.z, x, y, z)); }
        { const double4 q = sm.node4[q2]; d2n = sqrt(dist2(q.x, q.y, q.z, x, y, z)); }
        const bool pick1 = d1 < d2n;                     // trackdlo.cpp:324-329
        const int b = pick1 ? q1 : q2;
        const double db = pick1 ? d1 : d2n;
        const int lo = a < b ? a : b, hi = a < b ? b : a;
        const double dlo = a < b ? da : db, dhi = a < b ? db : da;
        // scaled geodesic coordinate: t = sqrt(0.5/sigma2) * (|s_j - s_ref| + d_ref)
        const double alo = sm.node4[lo].w + dlo * rscale;        //  s'_lo + d'_lo  (minus s'_j)
        const double ahi = dhi * rscale - sm.node4[hi].w;        //  d'_hi - s'_hi  (plus  s'_j)

        // ---- node window of this warp: outside [jlo, jhi] every P entry of these 32 points is EXACTLY 0 in
        // the reference (exp underflows for -0.5*geo/sigma2 < -745.13), so those rows are skipped.  A lane
        // needs j <= lo while s'_j > alo - T and j >= hi while s'_j < T - ahi (T^2 = 745.2); [lo, hi] itself
        // is always kept (the end quirk puts P = 1 between them).  The points are sorted by nearest node,
        // so the union over the warp stays narrow once sigma2 is small.
        int jlo, jhi;
        {
            const double T = 27.298351598585583;              // sqrt(745.2)
            double mlo = valid ? alo : 1e300, mhi = valid ? ahi : 1e300;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                mlo = fmin(mlo, __shfl_xor_sync(0xffffffffu, mlo, o));
                mhi = fmin(mhi, __shfl_xor_sync(0xffffffffu, mhi, o));
            }
            const int lomin = __reduce_min_sync(0xffffffffu, valid ? lo : Nn);
            const int himax = __reduce_max_sync(0xffffffffu, valid ? hi : -1);
            const double thr_lo = mlo - T, thr_hi = T - mhi;
            jlo = Nn; jhi = -1;
            for (int c = 0; c < Nn; c += 32) {
                const int j = c + lane;
                const double sj = j < Nn ? sm.node4[j].w : 0.0;
                const unsigned m1 = __ballot_sync(0xffffffffu, j < Nn && sj > thr_lo);
                const unsigned m2 = __ballot_sync(0xffffffffu, j < Nn && sj < thr_hi);
                if (m1 && jlo == Nn) jlo = c + __ffs(m1) - 1;
                if (m2) jhi = c + 31 - __clz(m2);
            }
            jlo = min(jlo, lomin); jhi = max(jhi, himax);
        }

        // ---- phase A: P column (trackdlo.cpp:332-354, 358-375).  Four nodes per trip: all shared-memory
        // loads of a trip precede its stores, so the four exp chains are independent and interleave.
        double colsum = 0.0;
        {
            const double4* nd = sm.node4;
            const double* vw = sm.vw;
            double* pc = pcol + jlo * RS;
            int j = jlo;
            for (; j + 3 <= jhi; j += 4) {
                const double s0 = nd[j].w, s1 = nd[j + 1].w, s2 = nd[j + 2].w, s3 = nd[j + 3].w;
                double v0 = 1.0, v1 = 1.0, v2 = 1.0, v3 = 1.0;
                if (VIS) { v0 = vw[j]; v1 = vw[j + 1]; v2 = vw[j + 2]; v3 = vw[j + 3]; }
                const double t0 = (j <= lo) ? (alo - s0) : (ahi + s0);
                const double t1 = (j + 1 <= lo) ? (alo - s1) : (ahi + s1);
                const double t2 = (j + 2 <= lo) ? (alo - s2) : (ahi + s2);
                const double t3 = (j + 3 <= lo) ? (alo - s3) : (ahi + s3);
                double p0 = exp_neg(t0 * t0, tab), p1 = exp_neg(t1 * t1, tab), p2 = exp_neg(t2 * t2, tab), p3 = exp_neg(t3 * t3, tab);
                if (VIS) { p0 *= v0; p1 *= v1; p2 *= v2; p3 *= v3; }
                colsum += (p0 + p1) + (p2 + p3);
                pc[0] = p0; pc[RS] = p1; pc[2 * RS] = p2; pc[3 * RS] = p3;
                pc += 4 * RS;
            }
            for (; j <= jhi; j++) {
                const double sj = nd[j].w;
                const double t = (j <= lo) ? (alo - sj) : (ahi + sj);
                double p = exp_neg(t * t, tab);
                if (VIS) p *= vw[j];
                colsum += p;
                *pc = p;
                pc += RS;
            }
        }
        if (hi - lo == 2) {                              // row strictly between lo and hi keeps geodesic 0 (end quirk)
            const int jb = lo + 1;
            const double pn = VIS ? sm.vw[jb] : 1.0;
            colsum += pn - pcol[jb * RS];
            pcol[jb * RS] = pn;
        }
        const double den = colsum + c_norm;              // trackdlo.cpp:379 / 382
        const double w = valid ? 1.0 / den : 0.0;
        sxx = fma(colsum * w, x * x + y * y + z * z, sxx);   // Pt1_n * |x_n|^2 (trackdlo.cpp:418)
        wb[lane] = make_double4(w, w * x, w * y, w * z);
        __syncwarp();

        // ---- phase B: lane = node; P1 / PX over this warp's 32 points (trackdlo.cpp:387-389).
        // Lanes whose node lies outside the window are masked off; a pass with no node inside is skipped.
#pragma unroll
        for (int ps = 0; ps < NPASS; ps++) {
            const int m = lane + 32 * ps;
            if (m >= jlo && m <= jhi) {
                const double* __restrict__ prow = pt + m * RS;
#pragma unroll 8
                for (int nn = 0; nn < 32; nn++) {
                    const double p = prow[nn];
                    const double4 w4 = wb[nn];
                    acc[ps][0] = fma(p, w4.x, acc[ps][0]); acc[ps][1] = fma(p, w4.y, acc[ps][1]);
                    acc[ps][2] = fma(p, w4.z, acc[ps][2]); acc[ps][3] = fma(p, w4.w, acc[ps][3]);
                }
            }
        }
        __syncwarp();
    }

    // ---- cross-warp reduction in a fixed order (deterministic)
    __syncthreads();
    double* __restrict__ racc = sm.ptile;                 // [nw][Nn][4]; the P slices are dead now
#pragma unroll
    for (int ps = 0; ps < NPASS; ps++) {
        const int m = lane + 32 * ps;
        if (m < Nn) {
            double* dst = racc + ((long long)warp * Nn + m) * 4;
            dst[0] = acc[ps][0]; dst[1] = acc[ps][1]; dst[2] = acc[ps][2]; dst[3] = acc[ps][3];
        }
    }
    __syncthreads();
    for (int i = tid; i < 4 * Nn; i += nt) {
        double v = 0.0;
        for (int w = 0; w < nw; w++) v += racc[w * 4 * Nn + i];
        __stcg(part_out + i, v);
    }
    const double sx = block_sum(sxx, sm.red);
    if (tid == 0) __stcg(part_out + 4 * Nn, sx);
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// Gauss-Jordan elimination with implicit partial (row) pivoting on the augmented system
// AB = [A | B] (n x (n+3), row stride ld).  Replaces completeOrthogonalDecomposition().solve
// (trackdlo.cpp:415) for the full-rank A of this path.  Block-cooperative; W -> wsol[n][3].
// Returns non-zero (uniform) if a zero / non-finite pivot was met.
// ------------------------------------------------------------------------------------------
__device__ int gj_solve(double* AB, int n, int ld, int* prow, int* used, double* rpiv_slot, double* wsol) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5, nt = blockDim.x;
    const int ncol = n + 3;
    for (int i = tid; i < n; i += nt) used[i] = 0;
    int bad = 0;
    __syncthreads();
    for (int k = 0; k < n; k++) {
        if (warp == 0) {
            double best = -1.0;
            int bi = 0x7fffffff;
            for (int i = lane; i < n; i += 32) {
                if (!used[i]) {
                    const double v = fabs(AB[(long long)i * ld + k]);
                    if (v > best || !(v == v)) { best = v; bi = i; }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (lane == 0) {
                if (bi == 0x7fffffff) bi = 0;
                prow[k] = bi;
                used[bi] = 1;
                *rpiv_slot = 1.0 / AB[(long long)bi * ld + k];
            }
        }
        __syncthreads();
        const int p = prow[k];
        const double rp = *rpiv_slot;
        if (!(fabs(rp) <= 1.79e308)) bad = 1;          // pivot 0 -> inf, NaN -> NaN
        const double* __restrict__ prowp = AB + (long long)p * ld;
        for (int i = warp; i < n; i += nw) {
            if (i == p) continue;
            double* __restrict__ row = AB + (long long)i * ld;
            const double f = row[k] * rp;
            for (int j = k + 1 + lane; j < ncol; j += 32) row[j] = fma(-f, prowp[j], row[j]);
        }
        __syncthreads();
    }
    for (int i = tid; i < 3 * n; i += nt) {
        const int k = i / 3, d = i - 3 * k;
        const int p = prow[k];
        wsol[i] = AB[(long long)p * ld + n + d] / AB[(long long)p * ld + k];
    }
    __syncthreads();
    return bad;
}

// ------------------------------------------------------------------------------------------
// Fast path of the solve for n <= 64 with [A | B] in shared memory: same Gauss-Jordan
// elimination with implicit partial pivoting, but ONE block barrier per elimination step.
// Warp 0 is the pivot warp: while the other warps apply step k to columns k+2.., it computes
// the updated column k+1 itself, selects the next pivot with three warp reductions on the bit
// patterns of |a_ik| (non-negative doubles order like their bits), and publishes the scaled
// multipliers for step k+1 (double-buffered).  gj: 2*64 multipliers + 1 flag.
// ------------------------------------------------------------------------------------------
// Pivot selection among the candidate rows {lane, lane+32}: arg-max of |v| over the warp via two
// integer reductions on the bit pattern (non-negative doubles order like their bits) + one vote.
// Ties go to the lowest lane.  NaN has the largest key, is selected, and poisons 1/pivot (-> flagged).
__device__ __forceinline__ void gj_pick(double v0, double v1, bool ok0, bool ok1, int lane, int& prow_out, double& pval_out) {
    const bool take1 = ok1 && (!ok0 || !(fabs(v1) <= fabs(v0)));
    const double val = take1 ? v1 : v0;
    const bool ok = ok0 || ok1;
    const int hi = ok ? (__double2hiint(val) & 0x7fffffff) : -1;
    const int mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned lo = (hi == mh) ? (unsigned)__double2loint(val) : 0u;
    const unsigned ml = __reduce_max_sync(0xffffffffu, lo);
    const unsigned match = __ballot_sync(0xffffffffu, ok && hi == mh && lo == ml);
    const unsigned t1 = __ballot_sync(0xffffffffu, take1);
    const int win = __ffs(match) - 1;
    prow_out = win + (((t1 >> win) & 1u) ? 32 : 0);
    pval_out = __shfl_sync(0xffffffffu, val, win);
}

// AB must point into shared memory (pass the __shared__-derived pointer directly so that the
// compiler emits LDS/STS with 32-bit addresses).
// pivot == false: the caller guarantees a symmetric positive definite A (elimination in natural order is
// stable, every pivot is positive); the search is skipped and the pivot warp's chain is ~3x shorter.
__device__ int gj_solve_small(double* __restrict__ AB, int n, int ld, double* __restrict__ gj, int* __restrict__ pivs,
                              double* __restrict__ wsol, const bool pivot, unsigned long long* dbg = nullptr) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5, nt = blockDim.x;
    const int ncol = n + 3;
    const int r0 = lane, r1 = lane + 32;
    const bool in0 = r0 < n, in1 = r1 < n;
    bool used0 = false, used1 = false;             // pivot warp: rows r0 / r1 already used as pivots
    int bad = 0;
    double* const a0p = AB + r0 * ld;              // pivot warp: its two rows
    double* const a1p = AB + r1 * ld;
    if (warp == 0) {
        const double v0 = in0 ? a0p[0] : 0.0, v1 = in1 ? a1p[0] : 0.0;
        int p; double pv;
        if (pivot) gj_pick(v0, v1, in0, in1, lane, p, pv);
        else { p = 0; pv = __shfl_sync(0xffffffffu, v0, 0); }
        const double rp = __drcp_rn(pv);
        if (!(fabs(rp) <= 1.79e308)) bad = 1;
        used0 |= (p == r0); used1 |= (p == r1);
        if (in0) gj[r0] = (r0 == p) ? 0.0 : v0 * rp;
        if (in1) gj[r1] = (r1 == p) ? 0.0 : v1 * rp;
        if (lane == 0) pivs[0] = p;
    }
    __syncthreads();
    // updater-warp geometry, hoisted out of the elimination loop (every instruction inside it is on the
    // critical path of a ~50-step dependent chain).  Phase "two" (more than 32 columns left): 2 column chunks x
    // RG2 row groups; phase "one": 1 chunk x RG1 row groups.
    const int U = nw - 1;
    const int RG2 = U >> 1, c2 = RG2 ? (warp - 1) / RG2 : 2, g2 = (warp - 1) - c2 * RG2;
    const int cnt2 = (warp >= 1 && RG2 && c2 <= 1 && g2 < n) ? (n - g2 + RG2 - 1) / RG2 : 0;
    const int RG1 = U > 0 ? U : 1, g1 = warp - 1;
    const int cnt1 = (g1 >= 0 && g1 < n) ? (n - g1 + RG1 - 1) / RG1 : 0;
    const int ktwo = (U >= 2) ? ncol - 34 : 0;                 // k < ktwo  <=>  ncol - (k + 2) > 32
    double* const pp2 = AB + g2 * ld + 2 + 32 * (c2 & 1) + lane;  // + k at step k
    double* const pp1 = AB + (g1 > 0 ? g1 : 0) * ld + 2 + lane;
    const int sr2 = RG2 * ld, sr1 = RG1 * ld;
    const int jo2 = 2 + 32 * (c2 & 1) + lane, jo1 = 2 + lane;   // column of this lane = k + jo
    for (int k = 0; k < n; k++) {
        const long long dbg_t0 = dbg ? clock64() : 0;
        const double* __restrict__ fk = gj + (k & 1) * 64;
        const int p = pivs[k];
        const double* __restrict__ prw = AB + p * ld;
        if (warp == 0) {
            // column k+1 belongs to the pivot warp in step k (for k = n-1 it is the first RHS column)
            const double pj = prw[k + 1];
            double v0 = 0.0, v1 = 0.0;
            if (in0) { v0 = fma(-fk[r0], pj, a0p[k + 1]); a0p[k + 1] = v0; }
            if (in1) { v1 = fma(-fk[r1], pj, a1p[k + 1]); a1p[k + 1] = v1; }
            if (k + 1 < n) {
                double* __restrict__ fn = gj + ((k + 1) & 1) * 64;
                int pn; double pv;
                if (pivot) gj_pick(v0, v1, in0 && !used0, in1 && !used1, lane, pn, pv);
                else { pn = k + 1; pv = __shfl_sync(0xffffffffu, pn < 32 ? v0 : v1, pn & 31); }
                const double rp = __drcp_rn(pv);
                bad |= !(fabs(rp) <= 1.79e308);
                used0 |= (pn == r0); used1 |= (pn == r1);
                // rows >= n hold v = 0; gj has 64 slots per buffer, so the stores need no guard
                fn[r0] = (r0 == pn) ? 0.0 : v0 * rp;
                fn[r1] = (r1 == pn) ? 0.0 : v1 * rp;
                pivs[k + 1] = pn;
            }
        } else {
            // updater warps 1..nw-1.  While more than 32 columns remain: 2 column chunks x (U/2) row groups,
            // afterwards 1 chunk x U row groups (U = nw-1 updater warps).  thread -> column j, rows g, g+RG, ...
            // Loads are issued in batches of 6 rows ahead of the FMAs/stores (latency-bound code).
            const bool two = k < ktwo;
            const int RG = two ? RG2 : RG1;
            const int c = two ? c2 : 0;
            const int g = two ? g2 : g1;
            const int sr = two ? sr2 : sr1;
            const int j = k + (two ? jo2 : jo1);
            int cnt = two ? cnt2 : cnt1;
            if (j < ncol && cnt > 0) {
                const double q = -prw[j];
                double* __restrict__ pp = (two ? pp2 : pp1) + k;
                const double* __restrict__ fp = fk + g;
                for (; cnt >= 6; cnt -= 6) {
                    const double x0 = pp[0], x1 = pp[sr], x2 = pp[2 * sr], x3 = pp[3 * sr], x4 = pp[4 * sr], x5 = pp[5 * sr];
                    const double f0 = fp[0], f1 = fp[RG], f2 = fp[2 * RG], f3 = fp[3 * RG], f4 = fp[4 * RG], f5 = fp[5 * RG];
                    pp[0] = fma(f0, q, x0); pp[sr] = fma(f1, q, x1); pp[2 * sr] = fma(f2, q, x2);
                    pp[3 * sr] = fma(f3, q, x3); pp[4 * sr] = fma(f4, q, x4); pp[5 * sr] = fma(f5, q, x5);
                    pp += 6 * sr; fp += 6 * RG;
                }
                if (cnt >= 3) {
                    const double x0 = pp[0], x1 = pp[sr], x2 = pp[2 * sr];
                    const double f0 = fp[0], f1 = fp[RG], f2 = fp[2 * RG];
                    pp[0] = fma(f0, q, x0); pp[sr] = fma(f1, q, x1); pp[2 * sr] = fma(f2, q, x2);
                    pp += 3 * sr; fp += 3 * RG; cnt -= 3;
                }
                for (; cnt > 0; cnt--) { pp[0] = fma(fp[0], q, pp[0]); pp += sr; fp += RG; }
            }
            if (n + 3 > 66 && c == (two ? 1 : 0)) {     // columns beyond the chunks above (n = 64, k = 0; or a single updater chunk)
                for (int j2 = k + 2 + (two ? 64 : 32) + lane; j2 < ncol; j2 += 32)
                    for (int i = g; i < n; i += RG) AB[i * ld + j2] = fma(-fk[i], prw[j2], AB[i * ld + j2]);
            }
        }
        if (dbg && lane == 0 && warp <= 1) atomicAdd(dbg + 14 + warp, (unsigned long long)(clock64() - dbg_t0));
        __syncthreads();
    }
    if (warp == 0 && lane == 0) gj[128] = (double)bad;
    for (int i = tid; i < 3 * n; i += nt) {
        const int k = i / 3, d = i - 3 * k;
        const int p = pivs[k];
        wsol[i] = AB[p * ld + n + d] / AB[p * ld + k];
    }
    __syncthreads();
    return gj[128] != 0.0;
}

// ------------------------------------------------------------------------------------------
// Register-resident Gauss-Jordan for n <= 64: [A | B] never lives in shared memory during the elimination
// (the shared-memory version above is bound by shared-memory bandwidth: every step re-reads and re-writes the
// whole matrix).  Thread (warp w, lane l) owns rows {l, l+32} x columns [w*CW, (w+1)*CW) in registers
// (needs nwarps * CW >= n + 3).  Per elimination step the owner warp of column k+1 updates that column first
// and publishes it (raw entries + the next pivot row / value) to a double-buffered shared array; everything
// else a thread needs is its own two multipliers (two LDS) and the pivot-row entries of its columns (warp
// shuffles from the lane that owns the pivot row).  ONE block barrier per step, no shared-memory matrix traffic.
// Implicit partial pivoting (PIVOT) or natural order (SPD systems).  Solution -> wsol[n][3].
// buf (doubles): mcol[2][64] @0 | pvv[2] @128 | pivots[64] @130 | out[64][3] @194 | ints: pivi[2], prow[64], flag @386
// ------------------------------------------------------------------------------------------
constexpr int GJR_BUF_DOUBLES = 386 + 40 + 16 * 12;     // + per-warp pivot-row buffers [<=16 warps][12]

__device__ __forceinline__ double rcp_fast(double x) {          // ~1 ulp reciprocal, no slow path (callers flag non-finite results)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);          // |e| <= 2^-20
    return fma(r, fma(e, e, e), r);            // r (1 + e + e^2): relative error e^3 <= 2^-60
}

template <int CW, bool PIVOT>
__device__ int gj_solve_regs(const double* AB, int n, int ld, double* buf, double* wsol) {   // no __restrict__: buf is written by other threads
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nt = blockDim.x, nwarp = nt >> 5;
    const int ncol = n + 3;
    double* mcol = buf;                 // [2][64]
    double* pvv = buf + 128;            // [2]
    double* pivots = buf + 130;         // [64]
    double* out = buf + 194;            // [64][3]
    int* ibuf = reinterpret_cast<int*>(buf + 386);
    int* pivi = ibuf;                   // [2]
    int* prow = ibuf + 2;               // [64]
    int* flag = ibuf + 66;
    double* rbuf = buf + 426 + w * 12;   // this warp's pivot-row buffer (16-byte aligned: buf is)
    const int r0 = lane, r1 = lane + 32;
    const int cbase = w * CW;
    double a0[CW], a1[CW];
#pragma unroll
    for (int c = 0; c < CW; c++) {
        const int col = cbase + c;
        a0[c] = (r0 < n && col < ncol) ? AB[r0 * ld + col] : 0.0;
        a1[c] = (r1 < n && col < ncol) ? AB[r1 * ld + col] : 0.0;
    }
    bool used0 = r0 >= n, used1 = r1 >= n;
    int bad = 0;
    if (tid == 0) *flag = 0;
    if (w == 0) {                       // publish column 0
        int pn; double pvn;
        if (PIVOT) gj_pick(a0[0], a1[0], !used0, !used1, lane, pn, pvn);
        else { pn = 0; pvn = __shfl_sync(0xffffffffu, a0[0], 0); }
        mcol[r0] = a0[0]; mcol[r1] = a1[0];
        if (lane == 0) { pivi[0] = pn; pvv[0] = pvn; pivots[0] = pvn; prow[0] = pn; }
    }
    __syncthreads();
    const int nkb = (n + CW - 1) / CW;
    for (int kb = 0; kb < nkb; kb++) {
#pragma unroll
        for (int kk = 0; kk < CW; kk++) {
            const int k = kb * CW + kk;
            if (k >= n) break;
            const int cur = k & 1, nxt = cur ^ 1;
            const int p = pivi[cur];
            const double pv = pvv[cur];
            double m0 = mcol[cur * 64 + r0], m1 = mcol[cur * 64 + r1];
            if (p == r0) { used0 = true; m0 = 0.0; }
            if (p == r1) { used1 = true; m1 = 0.0; }
            const double rd = -rcp_fast(pv);
            bad |= !(fabs(rd) <= 1.79e308);
            m0 *= rd; m1 *= rd;                                   // a[i][j] += (m_i * -1/pivot) * r_j
            const int pl = p & 31;
            const bool ph = p >= 32;
            const int co = (kk + 1 < CW) ? kk + 1 : 0;            // static: column k+1 inside its owner warp
            const int wo = (kk + 1 < CW) ? kb : kb + 1;           // owner warp of column k+1
            const bool active = (w > kb) || (w == kb && kk + 1 < CW) ;   // this warp still has columns > k
            // pivot-row entries of this warp's columns: the lane that holds row p publishes them to the warp's
            // shared buffer (4 predicated 16-byte stores), everybody reads them back with broadcast loads
            if (active) {
                if (lane == pl) {
#pragma unroll
                    for (int c = 0; c < CW; c += 2)
                        *reinterpret_cast<double2*>(rbuf + c) = ph ? make_double2(a1[c], a1[c + 1]) : make_double2(a0[c], a0[c + 1]);
                }
                __syncwarp();
                double rj[CW];
#pragma unroll
                for (int c = 0; c < CW; c += 2) { const double2 v = *reinterpret_cast<const double2*>(rbuf + c); rj[c] = v.x; rj[c + 1] = v.y; }
                // owner of column k+1: update it first and publish
                if (w == wo && k + 1 < n) {
                    a0[co] = fma(m0, rj[co], a0[co]); a1[co] = fma(m1, rj[co], a1[co]);
                    int pn; double pvn;
                    if (PIVOT) gj_pick(a0[co], a1[co], !used0, !used1, lane, pn, pvn);
                    else { pn = k + 1; pvn = __shfl_sync(0xffffffffu, pn < 32 ? a0[co] : a1[co], pn & 31); }
                    mcol[nxt * 64 + r0] = a0[co]; mcol[nxt * 64 + r1] = a1[co];
                    if (lane == 0) { pivi[nxt] = pn; pvv[nxt] = pvn; pivots[k + 1] = pvn; prow[k + 1] = pn; }
                }
                // the remaining columns > k of this warp
#pragma unroll
                for (int c = 0; c < CW; c++) {
                    if (c == co && k + 1 < n && w == wo) continue;         // done above
                    if (w == kb && c <= kk) continue;                      // columns <= k are finished
                    a0[c] = fma(m0, rj[c], a0[c]); a1[c] = fma(m1, rj[c], a1[c]);
                }
                __syncwarp();                                             // rbuf is rewritten in the next step
            }
            __syncthreads();
        }
    }
    // right-hand sides -> shared, solution off the pivot rows
#pragma unroll
    for (int c = 0; c < CW; c++) {
        const int col = cbase + c;
        if (col >= n && col < ncol) {
            if (r0 < n) out[r0 * 3 + (col - n)] = a0[c];
            if (r1 < n) out[r1 * 3 + (col - n)] = a1[c];
        }
    }
    if (bad) *flag = 1;
    __syncthreads();
    for (int i = tid; i < 3 * n; i += nt) {
        const int k = i / 3, d = i - 3 * k;
        wsol[i] = out[prow[k] * 3 + d] / pivots[k];
    }
    (void)nwarp;
    __syncthreads();
    return *flag;
}

// ------------------------------------------------------------------------------------------
// Blocked Cholesky solve for the SPD form of the M-step system at Nn > 64 (C5: Nn = 200), where [A|B] does not fit
// in registers or shared memory: A (lower triangle, row-major, row stride ld, in L2-resident global scratch) is
// factorised panel by panel (NB columns), LEFT-looking: a panel is first brought up to date against all previous
// panels with FP64 tensor-core MMAs (mma.sync.m8n8k4.f64: C[8x8] -= L[8x4] L^T[4x8], both operands streamed from L2,
// C in registers), then its diagonal block is factorised by one warp and the rows below by one thread per row.
// Then L z = b and L^T x = z for the three right-hand sides, column-oriented so that no cross-thread reduction is
// needed.  One CTA; rhs / solution in wsol[n][3] (shared).  Returns non-zero on a non-positive / non-finite pivot.
// work: >= n * (NB + 4) + n + 64 doubles of shared memory.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma_884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NB>
__device__ int chol_solve_blocked(double* A, int n, int ld, double* work, double* wsol) {
    constexpr int LDP = NB + 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = blockDim.x, nw = nt >> 5;
    double* Lp = work;                       // [n][LDP] panel (rows k0.. of columns k0..k0+nb)
    double* dinv = work + n * LDP;           // [n] reciprocal diagonal of L
    int* flag = reinterpret_cast<int*>(dinv + n);
    if (tid == 0) *flag = 0;
    __syncthreads();
    const int fr_ = lane >> 2, fk = lane & 3;                     // MMA fragment row / k index of this lane
    for (int k0 = 0; k0 < n; k0 += NB) {
        const int nb = min(NB, n - k0), R = n - k0;
        // (1) LEFT-LOOKING panel update with FP64 tensor-core MMAs:  panel = A[k0:, k0:k0+nb] - L[k0:, 0:k0] L[k0:k0+nb, 0:k0]^T.
        // Every warp owns 8-row tiles of the panel; the K loop streams both operands from L2 (no read-modify-write of
        // a trailing matrix: the loads of a tile are independent and stay in flight together).
        {
            constexpr int TR = 4;                                  // row tiles a warp advances together (they share the B fragment)
            const int ntile = (R + 7) >> 3;
            for (int g0 = warp * TR; g0 < ntile; g0 += nw * TR) {
                for (int ct = 0; ct < NB / 8; ct++) {              // 8-column tiles of the panel
                    const int bcol = ct * 8 + fr_;                 // panel column whose L row feeds the B fragment
                    const bool bok = bcol < nb;
                    const double* bro = A + (long long)(k0 + (bok ? bcol : 0)) * ld + fk;
                    const int cc = ct * 8 + 2 * fk;                // this lane's first C column inside the panel
                    double c0[TR], c1[TR];
                    const double* arow[TR];
                    bool rok[TR];
#pragma unroll
                    for (int t = 0; t < TR; t++) {
                        const int row = k0 + (g0 + t) * 8 + fr_;
                        rok[t] = (g0 + t < ntile) && row < n;
                        arow[t] = A + (long long)(rok[t] ? row : k0) * ld + fk;
                        c0[t] = (rok[t] && cc < nb) ? A[(long long)row * ld + k0 + cc] : 0.0;
                        c1[t] = (rok[t] && cc + 1 < nb) ? A[(long long)row * ld + k0 + cc + 1] : 0.0;
                    }
#pragma unroll 4
                    for (int kk = 0; kk < k0; kk += 4) {
                        const double bv = bok ? bro[kk] : 0.0;
#pragma unroll
                        for (int t = 0; t < TR; t++) {
                            const double av = rok[t] ? -arow[t][kk] : 0.0;
                            dmma_884(c0[t], c1[t], av, bv);
                        }
                    }
#pragma unroll
                    for (int t = 0; t < TR; t++) {
                        if (rok[t]) {
                            double* dst = Lp + ((g0 + t) * 8 + fr_) * LDP + cc;
                            dst[0] = cc < nb ? c0[t] : 0.0; dst[1] = cc + 1 < nb ? c1[t] : 0.0;
                        }
                    }
                }
            }
        }
        __syncthreads();
        // (2) diagonal block: unblocked Cholesky by warp 0 (lane = row)
        if (warp == 0) {
            for (int c = 0; c < nb; c++) {
                const double x = Lp[c * LDP + c];
                double rs;
                asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(rs) : "d"(x));
                {   // two Newton steps: rs <- rs (1.5 - 0.5 x rs^2)
                    const double hx = 0.5 * x;
                    rs = rs * fma(-hx * rs, rs, 1.5);
                    rs = rs * fma(-hx * rs, rs, 1.5);
                }
                if (!(x > 0.0) || !(fabs(rs) <= 1.79e308)) { if (lane == 0) *flag = 1; }
                const double lrc = (lane > c && lane < nb) ? Lp[lane * LDP + c] * rs : 0.0;
                __syncwarp();                                     // every lane has read Lp[c][c] before lane c overwrites it
                if (lane == c) { Lp[c * LDP + c] = x * rs; dinv[k0 + c] = rs; }
                if (lane > c && lane < nb) Lp[lane * LDP + c] = lrc;
                __syncwarp();
                // rank-1 update of the remaining lower triangle: row = lane, columns c+1..lane
                if (lane > c && lane < nb) {
                    for (int c2 = c + 1; c2 <= lane; c2++) Lp[lane * LDP + c2] = fma(-lrc, Lp[c2 * LDP + c], Lp[lane * LDP + c2]);
                }
                __syncwarp();
            }
        }
        __syncthreads();
        // (3) rows below the diagonal block: L21 = A21 L11^-T, one thread per row
        for (int r = nb + tid; r < R; r += nt) {
            double* row = Lp + r * LDP;
            for (int c = 0; c < nb; c++) {
                double v = row[c];
                for (int c2 = 0; c2 < c; c2++) v = fma(-row[c2], Lp[c * LDP + c2], v);
                row[c] = v * dinv[k0 + c];
            }
        }
        __syncthreads();
        // (4) L panel back to global (read by the later panels and by the substitutions)
        for (int idx = tid; idx < R * NB; idx += nt) {
            const int r = idx / NB, c = idx - r * NB;
            if (c < nb && c <= r) A[(long long)(k0 + r) * ld + k0 + c] = Lp[r * LDP + c];
        }
        __syncthreads();
    }
    // ---- forward substitution L z = b (in place in wsol), panel by panel
    for (int k0 = 0; k0 < n; k0 += NB) {
        const int nb = min(NB, n - k0), R = n - k0;
        for (int idx = tid; idx < R * NB; idx += nt) {
            const int r = idx / NB, c = idx - r * NB;
            Lp[r * LDP + c] = (c < nb && c <= r) ? A[(long long)(k0 + r) * ld + k0 + c] : 0.0;
        }
        __syncthreads();
        if (warp == 0) {                                          // diagonal block: lane = row, column oriented
            double b0 = 0.0, b1 = 0.0, b2 = 0.0;
            if (lane < nb) { b0 = wsol[(k0 + lane) * 3]; b1 = wsol[(k0 + lane) * 3 + 1]; b2 = wsol[(k0 + lane) * 3 + 2]; }
            for (int c = 0; c < nb; c++) {
                const double di = dinv[k0 + c];
                const double z0 = __shfl_sync(0xffffffffu, b0, c) * di, z1 = __shfl_sync(0xffffffffu, b1, c) * di, z2 = __shfl_sync(0xffffffffu, b2, c) * di;
                if (lane == c) { b0 = z0; b1 = z1; b2 = z2; }
                else if (lane > c && lane < nb) { const double l = Lp[lane * LDP + c]; b0 = fma(-l, z0, b0); b1 = fma(-l, z1, b1); b2 = fma(-l, z2, b2); }
            }
            if (lane < nb) { wsol[(k0 + lane) * 3] = b0; wsol[(k0 + lane) * 3 + 1] = b1; wsol[(k0 + lane) * 3 + 2] = b2; }
        }
        __syncthreads();
        for (int r = nb + tid; r < R; r += nt) {                  // rows below: b_r -= L21[r][:] z
            double b0 = wsol[(k0 + r) * 3], b1 = wsol[(k0 + r) * 3 + 1], b2 = wsol[(k0 + r) * 3 + 2];
            for (int c = 0; c < nb; c++) {
                const double l = Lp[r * LDP + c];
                b0 = fma(-l, wsol[(k0 + c) * 3], b0); b1 = fma(-l, wsol[(k0 + c) * 3 + 1], b1); b2 = fma(-l, wsol[(k0 + c) * 3 + 2], b2);
            }
            wsol[(k0 + r) * 3] = b0; wsol[(k0 + r) * 3 + 1] = b1; wsol[(k0 + r) * 3 + 2] = b2;
        }
        __syncthreads();
    }
    // ---- backward substitution L^T x = z, panels from the last to the first
    const int last = ((n - 1) / NB) * NB;
    for (int k0 = last; k0 >= 0; k0 -= NB) {
        const int nb = min(NB, n - k0);
        // diagonal block (rows/cols k0..k0+nb) -> shared
        for (int idx = tid; idx < nb * NB; idx += nt) {
            const int r = idx / NB, c = idx - r * NB;
            Lp[r * LDP + c] = (c < nb && c <= r) ? A[(long long)(k0 + r) * ld + k0 + c] : 0.0;
        }
        __syncthreads();
        if (warp == 0) {                                          // x_c for c = nb-1 .. 0; lane = row index i < c gets z_i -= L[c][i] x_c
            double b0 = 0.0, b1 = 0.0, b2 = 0.0;
            if (lane < nb) { b0 = wsol[(k0 + lane) * 3]; b1 = wsol[(k0 + lane) * 3 + 1]; b2 = wsol[(k0 + lane) * 3 + 2]; }
            for (int c = nb - 1; c >= 0; c--) {
                const double di = dinv[k0 + c];
                const double x0 = __shfl_sync(0xffffffffu, b0, c) * di, x1 = __shfl_sync(0xffffffffu, b1, c) * di, x2 = __shfl_sync(0xffffffffu, b2, c) * di;
                if (lane == c) { b0 = x0; b1 = x1; b2 = x2; }
                else if (lane < c) { const double l = Lp[c * LDP + lane]; b0 = fma(-l, x0, b0); b1 = fma(-l, x1, b1); b2 = fma(-l, x2, b2); }
            }
            if (lane < nb) { wsol[(k0 + lane) * 3] = b0; wsol[(k0 + lane) * 3 + 1] = b1; wsol[(k0 + lane) * 3 + 2] = b2; }
        }
        __syncthreads();
        for (int i = tid; i < k0; i += nt) {                      // rows above: z_i -= sum_{j in block} L[j][i] x_j
            double b0 = wsol[i * 3], b1 = wsol[i * 3 + 1], b2 = wsol[i * 3 + 2];
            for (int c = 0; c < nb; c++) {
                const double l = A[(long long)(k0 + c) * ld + i];
                b0 = fma(-l, wsol[(k0 + c) * 3], b0); b1 = fma(-l, wsol[(k0 + c) * 3 + 1], b1); b2 = fma(-l, wsol[(k0 + c) * 3 + 2], b2);
            }
            wsol[i * 3] = b0; wsol[i * 3 + 1] = b1; wsol[i * 3 + 2] = b2;
        }
        __syncthreads();
    }
    return *flag;
}

// ------------------------------------------------------------------------------------------
// LLE weights, one node per thread (trackdlo.cpp:92-159).  Mirrors the operation order of
// oracle/trackdlo_oracle.cpp::lle_weights_node with explicitly rounded mul/add/div so that both
// produce identical bits (the 6x6 Gram matrices are rank 3; their inverse is rounding noise).
// Writes row i of E = I - L into E[i*ldE + ...] (row must be pre-zeroed).
// ------------------------------------------------------------------------------------------
__device__ int lle_lu6(double* a, int nb, int* piv) {
    int sign = 1;
    for (int k = 0; k < nb; k++) {
        int p = k;
        double best = fabs(a[k * 6 + k]);
        for (int i = k + 1; i < nb; i++) {
            const double v = fabs(a[i * 6 + k]);
            if (v > best) { best = v; p = i; }
        }
        piv[k] = p;
        if (p != k) {
            for (int j = 0; j < nb; j++) { const double t = a[k * 6 + j]; a[k * 6 + j] = a[p * 6 + j]; a[p * 6 + j] = t; }
            sign = -sign;
        }
        const double pv = a[k * 6 + k];
        if (pv != 0.0) {
            for (int i = k + 1; i < nb; i++) {
                const double l = __ddiv_rn(a[i * 6 + k], pv);
                a[i * 6 + k] = l;
                for (int j = k + 1; j < nb; j++) a[i * 6 + j] = __dsub_rn(a[i * 6 + j], __dmul_rn(l, a[k * 6 + j]));
            }
        }
    }
    return sign;
}

__device__ void lle_row(const double* __restrict__ y0 /*[n][3]*/, int M, int i, double* Erow) {
    int nbr[6];
    int nb = 0;
    const int k = 3;
    if (i - k < 0) { for (int t = 0; t <= i + k && t < M; t++) if (t != i) nbr[nb++] = t; }
    else if (i + k >= M) { for (int t = i - k; t <= M - 1; t++) if (t != i) nbr[nb++] = t; }
    else { for (int t = i - k; t <= i + k; t++) if (t != i) nbr[nb++] = t; }
    double comp[6][3];
    for (int r = 0; r < nb; r++)
        for (int d = 0; d < 3; d++) comp[r][d] = __dsub_rn(y0[i * 3 + d], y0[nbr[r] * 3 + d]);
    double g[36], lu[36];
    for (int t = 0; t < 36; t++) g[t] = 0.0;
    for (int a = 0; a < nb; a++)
        for (int b = 0; b < nb; b++)
            g[a * 6 + b] = __dadd_rn(__dadd_rn(__dmul_rn(comp[a][0], comp[b][0]), __dmul_rn(comp[a][1], comp[b][1])),
                                     __dmul_rn(comp[a][2], comp[b][2]));
    int piv[6];
    for (int t = 0; t < 36; t++) lu[t] = g[t];
    const int sign = lle_lu6(lu, nb, piv);
    double det = (double)sign;
    for (int t = 0; t < nb; t++) det = __dmul_rn(det, lu[t * 6 + t]);
    if (!(det != 0.0)) {
        for (int t = 0; t < nb; t++) g[t * 6 + t] = __dadd_rn(g[t * 6 + t], 0.00001);
        for (int t = 0; t < 36; t++) lu[t] = g[t];
        lle_lu6(lu, nb, piv);
    }
    double rs[6], tot = 0.0;
    for (int r = 0; r < nb; r++) rs[r] = 0.0;
    // explicit inverse column by column; accumulate the row sums (Gi_inv * 1) in column order
    for (int c = 0; c < nb; c++) {
        double bcol[6];
        for (int r = 0; r < nb; r++) bcol[r] = (r == c) ? 1.0 : 0.0;
        for (int t = 0; t < nb; t++) { const int p = piv[t]; if (p != t) { const double x = bcol[t]; bcol[t] = bcol[p]; bcol[p] = x; } }
        for (int r = 1; r < nb; r++) {
            double s = bcol[r];
            for (int t = 0; t < r; t++) s = __dsub_rn(s, __dmul_rn(lu[r * 6 + t], bcol[t]));
            bcol[r] = s;
        }
        for (int r = nb - 1; r >= 0; r--) {
            double s = bcol[r];
            for (int t = r + 1; t < nb; t++) s = __dsub_rn(s, __dmul_rn(lu[r * 6 + t], bcol[t]));
            bcol[r] = __ddiv_rn(s, lu[r * 6 + r]);
        }
        for (int r = 0; r < nb; r++) rs[r] = __dadd_rn(rs[r], bcol[r]);
    }
    for (int r = 0; r < nb; r++) tot = __dadd_rn(tot, rs[r]);
    Erow[i] = 1.0;
    for (int r = 0; r < nb; r++) Erow[nbr[r]] = __dsub_rn(0.0, __ddiv_rn(rs[r], tot));
}

// ------------------------------------------------------------------------------------------
// traverse_euclidean (trackdlo.cpp:584-898) + helpers (utils.cpp:172-241); single thread.
// Semantics follow oracle/trackdlo_oracle.cpp (explicit bounds where the reference has UB).
// ------------------------------------------------------------------------------------------
struct V3 { double x, y, z; };
__device__ __forceinline__ V3 ld3(const double* g, int i) { return {g[i * 3], g[i * 3 + 1], g[i * 3 + 2]}; }
__device__ __forceinline__ double vdist(V3 a, V3 b) { return sqrt(dist2(a.x, a.y, a.z, b.x, b.y, b.z)); }

__device__ bool is_between(V3 x, V3 a, V3 b) {
    const double xs[3] = {x.x, x.y, x.z}, as[3] = {a.x, a.y, a.z}, bs[3] = {b.x, b.y, b.z};
    bool in_bound = true;
    for (int i = 0; i < 3; i++) {
        if (!(as[i] - 0.0001 <= xs[i] && xs[i] <= bs[i] + 0.0001) &&
            !(bs[i] - 0.0001 <= xs[i] && xs[i] <= as[i] + 0.0001)) in_bound = false;
    }
    return in_bound;
}

__device__ int line_sphere(V3 A, V3 B, V3 C, double radius, V3* out) {
    const double a = dist2(A.x, A.y, A.z, B.x, B.y, B.z);
    const double b = 2 * ((B.x - A.x) * (A.x - C.x) + (B.y - A.y) * (A.y - C.y) + (B.z - A.z) * (A.z - C.z));
    const double c = dist2(A.x, A.y, A.z, C.x, C.y, C.z) - radius * radius;
    const double delta = b * b - 4 * a * c;
    int cnt = 0;
    if (delta < 0) return 0;
    if (delta > 0) {
        const double sq = sqrt(delta);
        const double d1 = (-b + sq) / (2 * a), d2 = (-b - sq) / (2 * a);
        const V3 p1 = {A.x + d1 * (B.x - A.x), A.y + d1 * (B.y - A.y), A.z + d1 * (B.z - A.z)};
        const V3 p2 = {A.x + d2 * (B.x - A.x), A.y + d2 * (B.y - A.y), A.z + d2 * (B.z - A.z)};
        if (is_between(p1, A, B)) out[cnt++] = p1;
        if (is_between(p2, A, B)) out[cnt++] = p2;
    } else {
        const double d1 = -b / (2 * a);
        const V3 p1 = {A.x + d1 * (B.x - A.x), A.y + d1 * (B.y - A.y), A.z + d1 * (B.z - A.z)};
        if (is_between(p1, A, B)) out[cnt++] = p1;
    }
    return cnt;
}

// scans segments i = from, from+dir, ... while (dir>0 ? i+1 <= bound : i >= bound); returns yielding i or -1
__device__ int pursue(const double* guide, int from, int dir, int bound, V3& centre, double look) {
    for (int i = from; dir > 0 ? (i + 1 <= bound) : (i >= bound); i += dir) {
        const V3 A = ld3(guide, i), B = ld3(guide, i + dir);
        V3 xs[2];
        const int n = line_sphere(A, B, centre, look, xs);
        if (n == 0) continue;
        if (n == 1 && vdist(xs[0], B) > vdist(centre, B)) continue;
        V3 pick = xs[0];
        if (n == 2 && !(vdist(xs[0], B) <= vdist(xs[1], B))) pick = xs[1];
        centre = pick;
        return i;
    }
    return -1;
}

__device__ __forceinline__ void emit4(double* out, int& cnt, double idx, V3 p) {
    out[cnt * 4] = idx; out[cnt * 4 + 1] = p.x; out[cnt * 4 + 2] = p.y; out[cnt * 4 + 3] = p.z;
    cnt++;
}

// returns number of pairs written to out ([<= G+1][4]); *err |= bits on reference-UB paths
__device__ int traverse_euclidean(const double* geo, int G, const double* guide, int R, const int* vis, int V,
                                  int alignment, int align_idx, double* out, int* err) {
    int cnt = 0;
    if (R == 1) { emit4(out, cnt, vis[0], ld3(guide, 0)); return cnt; }
    if (alignment == 0) {
        emit4(out, cnt, vis[0], ld3(guide, 0));
        int cs = 0;
        for (int i = 0; i < V; i++) { if (i == vis[i]) cs++; else break; }
        if (cs == 0) { *err |= 1; return cnt; }
        int last = 0, k = 0;
        V3 centre = ld3(guide, 0);
        while (last + 1 <= cs - 1 && k + 1 <= G - 1) {
            const double look = fabs(geo[k + 1] - geo[k]);
            const int got = pursue(guide, last, +1, cs - 1, centre, look);
            if (got < 0) break;
            last = got;
            emit4(out, cnt, k + 1, centre);
            k++;
        }
    } else if (alignment == 1) {
        emit4(out, cnt, vis[V - 1], ld3(guide, R - 1));
        int cs = 0;
        for (int i = 1; i <= V; i++) { if (vis[V - i] == G - i) cs++; else break; }
        int last = R - 1, k = G - 1;
        V3 centre = ld3(guide, R - 1);
        const int lowest = R - cs;
        while (last - 1 >= lowest && k - 1 >= 0) {
            const double look = fabs(geo[k] - geo[k - 1]);
            const int got = pursue(guide, last, -1, lowest + 1, centre, look);
            if (got < 0) break;
            last = got;
            emit4(out, cnt, k - 1, centre);
            k--;
        }
    } else {
        if (align_idx < 0 || align_idx >= V || align_idx >= R) { *err |= 2; return cnt; }
        emit4(out, cnt, vis[align_idx], ld3(guide, align_idx));
        int cs2 = 1;
        for (int i = align_idx + 1; i < V; i++) { if (vis[i] - vis[i - 1] == 1) cs2++; else break; }
        int last = align_idx, k = vis[align_idx];
        V3 centre = ld3(guide, align_idx);
        while (last + 1 <= align_idx + cs2 - 1 && k + 1 <= G - 1) {
            const double look = fabs(geo[k + 1] - geo[k]);
            const int got = pursue(guide, last, +1, align_idx + cs2 - 1, centre, look);
            if (got < 0) break;
            last = got;
            emit4(out, cnt, k + 1, centre);
            k++;
        }
        int cs1 = 1;
        if (align_idx - 1 >= 0)
            for (int i = align_idx - 1; i >= 0 && i + 1 < V; i++) { if (vis[i + 1] - vis[i] == 1) cs1++; else break; }
        last = align_idx; k = vis[align_idx];
        centre = ld3(guide, align_idx);
        while ((unsigned long long)(long long)(last - 1) >= (unsigned long long)(long long)align_idx - (unsigned long long)cs1
               && k - 1 >= 0) {
            const double look = fabs(geo[k] - geo[k - 1]);
            const int got = pursue(guide, last, -1, 1, centre, look);
            if (got < 0) break;
            last = got;
            emit4(out, cnt, k - 1, centre);
            k--;
        }
    }
    return cnt;
}

// ------------------------------------------------------------------------------------------
// One cpd_lle call executed by the whole cluster (trackdlo.cpp:161-441).
// All CTAs execute the same sequence of cluster barriers.  Returns the status mask (uniform).
// Yio: global [Nn][3] in/out.  sigma2_out / Wout / iters_out may be null.
// ------------------------------------------------------------------------------------------
template <int NPASS>
__device__ int cpd_run(cg::cluster_group& cluster, const Smem& sm, const KArgs& a, double* cscr,
                       const double* __restrict__ Xraw, long long m0, double* __restrict__ Xc,
                       double* Yio, int Nn, double sigma2_in, double* sigma2_out, const CpdP& p,
                       const double* priors, int n_priors, int n_visible, const double* Hext, int hstride,
                       double* Wout, int* iters_out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5, nt = blockDim.x;
    const int rank = (int)cluster.block_rank();
    const int C = (int)cluster.num_blocks();
    const Scr sc = scr_layout(a.scr_nodes);
    double* gG = cscr + sc.G;
    double* gHG = cscr + sc.HG;
    double* gH = cscr + sc.H;
    double* gPART = cscr + sc.PART;
    double* gDMIN = cscr + sc.DMIN;
    double* gGATH = cscr + sc.GATH;
    double* gSTATE = cscr + sc.STATE;
    const int ld = Nn + 3;
    const bool ab_in_smem = (long long)Nn * (Nn + 3) <= (long long)(a.tile / 32) * a.nmax * 33;
    double* AB = ab_in_smem ? sm.ptile : (cscr + sc.AB);

    if (Nn < 4) {                         // reference indexes rows 2 and Nn-3 (trackdlo.cpp:313-321)
        if (rank == 0 && tid == 0) { if (iters_out) *iters_out = 0; }
        return ST_TOO_FEW_NODES;
    }

    // ---- Y -> shared (Y0 and current Y), arc-length coordinates s (trackdlo.cpp:203, 216-223)
    for (int i = tid; i < 3 * Nn; i += nt) sm.y0[i] = Yio[i];
    __syncthreads();
    for (int j = tid; j < Nn; j += nt) sm.node4[j] = make_double4(sm.y0[3 * j], sm.y0[3 * j + 1], sm.y0[3 * j + 2], 0.0);
    if (tid == 0) {
        double cur = 0.0;
        sm.s[0] = 0.0;
        for (int i = 0; i < Nn - 1; i++) {
            cur += sqrt(dist2(sm.y0[3 * i + 3], sm.y0[3 * i + 4], sm.y0[3 * i + 5], sm.y0[3 * i], sm.y0[3 * i + 1], sm.y0[3 * i + 2]));
            sm.s[i + 1] = cur;
        }
    }
    __syncthreads();

    // ---- prune + compaction of this CTA's slice (trackdlo.cpp:177-195)
    const long long per = (m0 + C - 1) / C;
    long long r0 = per * rank, r1 = r0 + per;
    if (r0 > m0) r0 = m0;
    if (r1 > m0) r1 = m0;
    double sum_local;
    const int n_local = prune_sort_slice(sm, Xraw, r0, r1, Xc, a.bkt + (Xraw - a.X) / 3, Nn, p.prune_radius, &sum_local);
    if (tid == 0) { __stcg(gGATH + 2 * rank, (double)n_local); __stcg(gGATH + 2 * rank + 1, sum_local); }
    const double* Xloc = Xc + r0 * 3;

    // ---- rank 0: G, LLE products, priors (trackdlo.cpp:225-260)
    if (rank == 0) {
        const double beta = p.beta;
        for (int idx = tid; idx < Nn * Nn; idx += nt) {
            const int i = idx / Nn, j = idx - i * Nn;
            const double dd = fabs(sm.s[i] - sm.s[j]);
            gG[idx] = 1 / (2 * beta * 2 * beta) * exp(-sqrt(2.0) * dd / beta) * (2 * dd + sqrt(2.0) * beta);
        }
        for (int i = tid; i < Nn; i += nt) sm.jd[i] = 0.0;
        for (int i = tid; i < 3 * Nn; i += nt) sm.yext[i] = sm.y0[i];
        __syncthreads();
        if (tid == 0) {
            for (int k = 0; k < n_priors; k++) {
                const int idx = (int)priors[k * 4];
                if (idx < 0 || idx >= Nn) continue;
                sm.jd[idx] = 1.0;
                sm.yext[idx * 3] = priors[k * 4 + 1]; sm.yext[idx * 3 + 1] = priors[k * 4 + 2]; sm.yext[idx * 3 + 2] = priors[k * 4 + 3];
            }
        }
        if (p.include_lle) {
            if (Hext) {
                for (int idx = tid; idx < Nn * Nn; idx += nt) { const int i = idx / Nn, j = idx - i * Nn; gH[idx] = Hext[(long long)i * hstride + j]; }
            } else {
                double* E = cscr + sc.AB;                              // dense E = I - L
                for (int idx = tid; idx < Nn * Nn; idx += nt) E[idx] = 0.0;
                __syncthreads();
                for (int i = tid; i < Nn; i += nt) lle_row(sm.y0, Nn, i, E + (long long)i * Nn);
                __syncthreads();
                for (int idx = tid; idx < Nn * Nn; idx += nt) {        // H = E^T E, k ascending, band only
                    const int r = idx / Nn, c = idx - r * Nn;
                    double s = 0.0;
                    int klo = (r > c ? r : c) - 3, khi = (r < c ? r : c) + 3;
                    if (klo < 0) klo = 0;
                    if (khi > Nn - 1) khi = Nn - 1;
                    for (int k = klo; k <= khi; k++) s = __dadd_rn(s, __dmul_rn(E[(long long)k * Nn + r], E[(long long)k * Nn + c]));
                    gH[idx] = s;
                }
            }
            __syncthreads();
            for (int idx = tid; idx < Nn * Nn; idx += nt) {            // HG = H G
                const int i = idx / Nn, j = idx - i * Nn;
                double s = 0.0;
                for (int k = 0; k < Nn; k++) s = fma(gH[(long long)i * Nn + k], gG[(long long)k * Nn + j], s);
                gHG[idx] = s;
            }
            for (int idx = tid; idx < 3 * Nn; idx += nt) {             // H Y0
                const int i = idx / 3, d = idx - 3 * i;
                double s = 0.0;
                for (int k = 0; k < Nn; k++) s = fma(gH[(long long)i * Nn + k], sm.y0[3 * k + d], s);
                sm.hy0[idx] = s;
            }
        }
        __syncthreads();
    }

    cluster.sync();                                                    // (S) gather prune results
    long long Mp = 0;
    double sumd2 = 0.0;
    for (int r = 0; r < C; r++) { Mp += (long long)__ldcg(gGATH + 2 * r); sumd2 += __ldcg(gGATH + 2 * r + 1); }
    if (Mp == 0) {
        if (rank == 0 && tid == 0) { if (iters_out) *iters_out = 0; }
        cluster.sync();
        return ST_EMPTY;
    }
    double sigma2 = sigma2_in;
    if (sigma2 == 0) sigma2 = sumd2 / (3.0 * (double)Nn * (double)Mp);   // trackdlo.cpp:271-273
    const bool use_vis = (n_visible != Nn) && (n_visible > 0) && (p.k_vis != 0);   // trackdlo.cpp:358
    const bool have_priors = n_priors > 0;

    int status = 0, iters = 0;
    // optional phase timers (thread 0 of every CTA): 0 setup, 1 dmin, 2 estep, 3 wait after estep, 4 mstep, 5 wait after mstep
    long long tprev = a.prof ? clock64() : 0;
#define TDLO_TICK(slot)                                                                                   \
    if (a.prof && tid == 0) { const long long tn_ = clock64(); atomicAdd(a.prof + (rank ? 8 : 0) + (slot), (unsigned long long)(tn_ - tprev)); tprev = tn_; }
    for (int it = 0; it < p.max_iter; it++) {
        iters = it + 1;
        const double rscale = sqrt(0.5 / sigma2);
        for (int j = tid; j < Nn; j += nt) sm.node4[j].w = sm.s[j] * rscale;
        const double c_gauss = pow(2 * M_PI * sigma2, 1.5) * p.mu / (1 - p.mu);
        double c_norm = c_gauss * (double)Nn / (double)Mp;             // trackdlo.cpp:300
        __syncthreads();
        TDLO_TICK(0)
        if (use_vis) {
            dmin_slice(sm, Xloc, n_local, Nn, gDMIN + rank * Nn);
            cluster.sync();                                            // (V)
            for (int j = tid; j < Nn; j += nt) {
                double m = __ldcg(gDMIN + j);
                for (int r = 1; r < C; r++) m = fmin(m, __ldcg(gDMIN + r * Nn + j));
                double dm = sqrt(m);
                if (dm <= p.tau) dm = 0.0;                              // trackdlo.cpp:291-293
                sm.vw[j] = exp(-p.k_vis * dm);                          // trackdlo.cpp:365
            }
            __syncthreads();
            if (tid == 0) { double t = 0.0; for (int j = 0; j < Nn; j++) t += sm.vw[j]; sm.red[40] = t; }
            __syncthreads();
            const double tot = sm.red[40];
            for (int j = tid; j < Nn; j += nt) sm.vw[j] = sm.vw[j] / tot;   // trackdlo.cpp:372
            c_norm = c_gauss / (double)Mp;                              // trackdlo.cpp:378
            __syncthreads();
            TDLO_TICK(1)
            estep_slice<NPASS, true>(sm, Xloc, n_local, Nn, sigma2, c_norm, rscale, gPART + rank * (4 * Nn + 4));
        } else {
            estep_slice<NPASS, false>(sm, Xloc, n_local, Nn, sigma2, c_norm, rscale, gPART + rank * (4 * Nn + 4));
        }
        TDLO_TICK(2)
        cluster.sync();                                                // (1) partial sums visible
        TDLO_TICK(3)

        if (rank == 0) {
            // ---- gather partials in rank order
            for (int i = tid; i < 4 * Nn; i += nt) {
                double v = 0.0;
                for (int r = 0; r < C; r++) v += __ldcg(gPART + r * (4 * Nn + 4) + i);
                const int m = i >> 2, k = i & 3;
                if (k == 0) sm.p1[m] = v; else sm.px[3 * m + k - 1] = v;
            }
            double sxx = 0.0;
            for (int r = 0; r < C; r++) sxx += __ldcg(gPART + r * (4 * Nn + 4) + 4 * Nn);
            __syncthreads();
            // ---- assemble [A | B] (trackdlo.cpp:392-413)
            const double ls = p.lambda * sigma2, sg = sigma2 * p.gamma;
            // Without LLE, A = D G + ls I with D = diag(P1 + alpha J) is similar to the SPD matrix
            // D^1/2 G D^1/2 + ls I (G is a Matern-3/2 kernel of the arc length): solve that one without
            // pivoting, (D^1/2 G D^1/2 + ls I) Z = D^-1/2 B, W = D^1/2 Z.  Rows with D_i = 0 have B_i = 0 and give
            // W_i = 0 in both forms.  Same conditioning as A.
            const bool spd = !p.include_lle && ab_in_smem && Nn <= 64 && nt >= 64;
            if (spd) {
                for (int i = tid; i < Nn; i += nt) sm.tnew[i] = sqrt(sm.p1[i] + (have_priors ? p.alpha * sm.jd[i] : 0.0));
                __syncthreads();
            }
            for (int idx = tid; idx < Nn * Nn; idx += nt) {
                const int i = idx / Nn, j = idx - i * Nn;
                const double g = gG[idx];
                double v;
                if (spd) v = sm.tnew[i] * g * sm.tnew[j] + (i == j ? ls : 0.0);
                else {
                    v = sm.p1[i] * g + (i == j ? ls : 0.0);
                    if (p.include_lle) v += sg * gHG[idx];
                    if (have_priors) v += p.alpha * sm.jd[i] * g;
                }
                AB[(long long)i * ld + j] = v;
            }
            for (int idx = tid; idx < 3 * Nn; idx += nt) {
                const int i = idx / 3, d = idx - 3 * i;
                double v = sm.px[idx] - sm.p1[i] * sm.y0[idx];
                if (p.include_lle) v -= sg * sm.hy0[idx];
                if (have_priors) v += p.alpha * (sm.yext[idx] - sm.y0[idx]);
                if (spd) { const double sd = sm.tnew[i]; v = sd > 0.0 ? v / sd : 0.0; }
                AB[(long long)i * ld + Nn + d] = v;
            }
            __syncthreads();
            TDLO_TICK(4)
            int sing;
            if (ab_in_smem && Nn <= 64 && nt >= 64) {
                double sdreg[3];
                if (spd) for (int t = 0, i = tid; t < 3 && i < 3 * Nn; t++, i += nt) sdreg[t] = sm.tnew[i / 3];   // tnew is reused below
                sing = gj_solve_small(sm.ptile, Nn, ld, sm.gjbuf, sm.prow, sm.wsol, !spd, a.prof);
                if (spd) {
                    for (int t = 0, i = tid; t < 3 && i < 3 * Nn; t++, i += nt) sm.wsol[i] *= sdreg[t];
                    __syncthreads();
                }
            } else sing = gj_solve(AB, Nn, ld, sm.prow, sm.used, sm.red + 41, sm.wsol);
            if (sing) status |= ST_SINGULAR;
            TDLO_TICK(6)
            // ---- T = Y0 + G W (trackdlo.cpp:417)
            for (int i = warp; i < Nn; i += nw) {
                double ax = 0.0, ay = 0.0, az = 0.0;
                for (int k = lane; k < Nn; k += 32) {
                    const double g = gG[(long long)i * Nn + k];
                    ax = fma(g, sm.wsol[3 * k], ax); ay = fma(g, sm.wsol[3 * k + 1], ay); az = fma(g, sm.wsol[3 * k + 2], az);
                }
                ax = warp_sum(ax); ay = warp_sum(ay); az = warp_sum(az);
                if (lane == 0) { sm.tnew[3 * i] = sm.y0[3 * i] + ax; sm.tnew[3 * i + 1] = sm.y0[3 * i + 1] + ay; sm.tnew[3 * i + 2] = sm.y0[3 * i + 2] + az; }
            }
            __syncthreads();
            // ---- sigma2 update and convergence test (trackdlo.cpp:418-431)
            if (warp == 0) {
                double np = 0.0, trPXT = 0.0, trTPT = 0.0, moved = 0.0;
                for (int m = lane; m < Nn; m += 32) {
                    const double tx = sm.tnew[3 * m], ty = sm.tnew[3 * m + 1], tz = sm.tnew[3 * m + 2];
                    const double p1 = sm.p1[m];
                    np += p1;
                    trPXT += sm.px[3 * m] * tx + sm.px[3 * m + 1] * ty + sm.px[3 * m + 2] * tz;
                    trTPT += p1 * (tx * tx + ty * ty + tz * tz);
                    const double4 yc = sm.node4[m];
                    moved += sqrt(dist2(yc.x, yc.y, yc.z, tx, ty, tz));
                }
                np = warp_sum(np); trPXT = warp_sum(trPXT); trTPT = warp_sum(trTPT); moved = warp_sum(moved);
                if (lane == 0) {
                    const double s2new = (sxx - 2 * trPXT + trTPT) / (np * 3);
                    const bool done = (moved / Nn) < p.tol;
                    int fin = 0;
                    if (done) fin = 1;
                    else if (it == p.max_iter - 1) { fin = 1; status |= ST_NOT_CONVERGED; }
                    __stcg(gSTATE + 3 * Nn, s2new);
                    __stcg(gSTATE + 3 * Nn + 1, (double)fin);
                    __stcg(gSTATE + 3 * Nn + 2, (double)status);
                }
            }
            for (int i = tid; i < 3 * Nn; i += nt) __stcg(gSTATE + i, sm.tnew[i]);
        }
        TDLO_TICK(7)
        cluster.sync();                                                // (2) new state visible
        TDLO_TICK(5)
        for (int j = tid; j < Nn; j += nt) {
            const double s = sm.node4[j].w;
            sm.node4[j] = make_double4(__ldcg(gSTATE + 3 * j), __ldcg(gSTATE + 3 * j + 1), __ldcg(gSTATE + 3 * j + 2), s);
        }
        sigma2 = __ldcg(gSTATE + 3 * Nn);
        const int fin = (int)__ldcg(gSTATE + 3 * Nn + 1);
        status = (int)__ldcg(gSTATE + 3 * Nn + 2);
        __syncthreads();
        if (fin) break;
    }

#undef TDLO_TICK
    // ---- results (rank 0)
    if (rank == 0) {
        for (int j = tid; j < Nn; j += nt) {
            const double4 q = sm.node4[j];
            Yio[3 * j] = q.x; Yio[3 * j + 1] = q.y; Yio[3 * j + 2] = q.z;
        }
        if (Wout && p.max_iter > 0) for (int i = tid; i < 3 * Nn; i += nt) Wout[i] = sm.wsol[i];
        if (tid == 0) {
            if (sigma2_out) *sigma2_out = sigma2;
            if (iters_out) *iters_out = iters;
        }
    }
    // make Yio visible to the whole cluster (tracking mode reads it back) and protect scratch reuse
    __threadfence();
    cluster.sync();
    return status;
}

// ------------------------------------------------------------------------------------------
// The persistent kernel.
// ------------------------------------------------------------------------------------------
template <int NPASS, int MINB>
__global__ void __launch_bounds__(kMaxThreads, MINB) tdlo_em_kernel(const KArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x, nt = blockDim.x;
    const int rank = (int)cluster.block_rank();
    const int C = (int)cluster.num_blocks();
    const int cluster_id = blockIdx.x / C;

    const SmemL& L = a.L;
    Smem sm;
    sm.tab = reinterpret_cast<double*>(smem_raw + L.tab);
    sm.node4 = reinterpret_cast<double4*>(smem_raw + L.node4);
    sm.wbuf = reinterpret_cast<double4*>(smem_raw + L.wbuf);
    sm.y0 = reinterpret_cast<double*>(smem_raw + L.y0);
    sm.s = reinterpret_cast<double*>(smem_raw + L.s);
    sm.vw = reinterpret_cast<double*>(smem_raw + L.vw);
    sm.yext = reinterpret_cast<double*>(smem_raw + L.yext);
    sm.jd = reinterpret_cast<double*>(smem_raw + L.jd);
    sm.hy0 = reinterpret_cast<double*>(smem_raw + L.hy0);
    sm.p1 = reinterpret_cast<double*>(smem_raw + L.p1);
    sm.px = reinterpret_cast<double*>(smem_raw + L.px);
    sm.wsol = reinterpret_cast<double*>(smem_raw + L.wsol);
    sm.tnew = reinterpret_cast<double*>(smem_raw + L.tnew);
    sm.pacc = reinterpret_cast<double*>(smem_raw + L.pacc);
    sm.red = reinterpret_cast<double*>(smem_raw + L.red);
    sm.gjbuf = reinterpret_cast<double*>(smem_raw + L.gjbuf);
    sm.prow = reinterpret_cast<int*>(smem_raw + L.prow);
    sm.used = reinterpret_cast<int*>(smem_raw + L.used);
    sm.ptile = reinterpret_cast<double*>(smem_raw + L.ptile);

    for (int i = tid; i < 64; i += nt) sm.tab[i] = c_exp_tab[i];
    __syncthreads();

    double* cscr = a.scratch + (long long)cluster_id * a.scratch_stride;
    const Scr sc = scr_layout(a.scr_nodes);
    int* ctl = reinterpret_cast<int*>(cscr + sc.CTL);

    for (;;) {
        if (rank == 0 && tid == 0) __stcg(ctl, atomicAdd(a.queue, 1));
        cluster.sync();
        const int f = __ldcg(ctl);
        cluster.sync();
        if (f >= a.n_frames) break;

        const long long x0 = a.x_off[f], m0 = a.x_off[f + 1] - x0;
        const double* Xraw = a.X + x0 * 3;
        double* Xc = a.Xc + x0 * 3;

        if (a.mode == 0) {
            const int Nn = a.n_nodes ? a.n_nodes[f] : a.node_stride;
            const long long ys = (long long)f * a.node_stride;
            const int st = cpd_run<NPASS>(cluster, sm, a, cscr, Xraw, m0, Xc, a.Y + ys * 3, Nn, a.sigma2[f], a.sigma2 + f, a.p0,
                                        a.priors ? a.priors + ys * 4 : nullptr,
                                        (a.priors && a.n_priors) ? a.n_priors[f] : 0,
                                        a.n_visible ? a.n_visible[f] : 0,
                                        a.H ? a.H + ys * a.node_stride : nullptr, a.node_stride,
                                        a.W ? a.W + ys * 3 : nullptr, a.iters ? a.iters + f : nullptr);
            if (rank == 0 && tid == 0 && a.status) a.status[f] = st;
        } else {
            // ---------------- tracking_step (trackdlo.cpp:900-999)
            const int Nn = a.node_stride;
            double* Yf = a.Y + (long long)f * Nn * 3;
            const int* vis = a.vis + a.vis_off[f];
            const int nvis = (int)(a.vis_off[f + 1] - a.vis_off[f]);
            const int* ext = a.ext + a.ext_off[f];
            const int V = (int)(a.ext_off[f + 1] - a.ext_off[f]);
            double* guide = a.guide_out ? a.guide_out + (long long)f * Nn * 3 : cscr + sc.GUIDE;
            double* pri = a.priors_out ? a.priors_out + (long long)f * 2 * Nn * 4 : cscr + sc.PRI;
            const double* geo = a.rest + (long long)f * Nn;
            int st = 0;
            // guide nodes (trackdlo.cpp:913-921)
            if (rank == 0) {
                for (int i = tid; i < 3 * V; i += nt) {
                    const int r = i / 3, d = i - 3 * r;
                    guide[i] = (V != Nn) ? Yf[ext[r] * 3 + d] : Yf[i];
                }
                __threadfence();
            }
            cluster.sync();
            // pre-processing registration (trackdlo.cpp:925-927); sigma2 copy is discarded
            const int st_pre = cpd_run<NPASS>(cluster, sm, a, cscr, Xraw, m0, Xc, guide, V, a.sigma2[f], nullptr, a.p0,
                                            nullptr, 0, 0, a.H ? a.H + (long long)f * Nn * Nn : nullptr, Nn,
                                            nullptr, a.iters ? a.iters + 2 * f : nullptr);
            if (st_pre & ST_NOT_CONVERGED) st |= ST_PRE_NOT_CONVERGED;
            st |= st_pre & ~ST_NOT_CONVERGED;
            int* ictl = ctl + 2;
            if (rank == 0 && tid == 0) {
                int err = 0, state = 0, np = 0;
                double* trv = cscr + sc.TRV;
                if (st_pre & (ST_TOO_FEW_NODES | ST_EMPTY)) {
                    state = -1;
                } else if (V == Nn) {
                    state = 0;
                    double* v1 = trv;
                    double* v2 = trv + (Nn + 2) * 4;
                    const int n1 = traverse_euclidean(geo, Nn, guide, V, ext, V, 0, -1, v1, &err);
                    const int n2 = traverse_euclidean(geo, Nn, guide, V, ext, V, 1, -1, v2, &err);
                    // v2 is emitted tail -> head; the reference reverses it (trackdlo.cpp:942): v2r[j] = v2[n2-1-j]
                    for (int i = 0; i < Nn; i++) {
                        const int j2 = i - (Nn - n2);
                        const double* first2 = v2 + (n2 - 1) * 4;
                        if (i < first2[0] && i < n1) { for (int t = 0; t < 4; t++) pri[np * 4 + t] = v1[i * 4 + t]; np++; }
                        else if (i > v1[(n1 - 1) * 4] && j2 >= 0 && j2 < n2) {
                            const double* s2 = v2 + (n2 - 1 - j2) * 4;
                            for (int t = 0; t < 4; t++) pri[np * 4 + t] = s2[t];
                            np++;
                        } else if (i < n1 && j2 >= 0 && j2 < n2) {
                            const double* s2 = v2 + (n2 - 1 - j2) * 4;
                            for (int t = 0; t < 4; t++) pri[np * 4 + t] = (v1[i * 4 + t] + s2[t]) / 2.0;
                            np++;
                        } else err |= 4;
                    }
                } else if (ext[0] == 0 && ext[V - 1] == Nn - 1) {
                    state = 1;
                    np = traverse_euclidean(geo, Nn, guide, V, ext, V, 0, -1, pri, &err);
                    np += traverse_euclidean(geo, Nn, guide, V, ext, V, 1, -1, pri + np * 4, &err);
                } else if (ext[0] == 0) {
                    state = 2;
                    np = traverse_euclidean(geo, Nn, guide, V, ext, V, 0, -1, pri, &err);
                } else if (ext[V - 1] == Nn - 1) {
                    state = 3;
                    np = traverse_euclidean(geo, Nn, guide, V, ext, V, 1, -1, pri, &err);
                } else {
                    state = 4;
                    int align = -1;
                    double moved = 999999;
                    for (int i = 0; i < nvis; i++) {
                        if (i >= V) { err |= 8; break; }
                        const double dd = vdist(ld3(Yf, vis[i]), ld3(guide, i));
                        if (dd < moved) { moved = dd; align = i; }
                    }
                    np = traverse_euclidean(geo, Nn, guide, V, ext, V, 2, align, pri, &err);
                }
                if (a.state_out) a.state_out[f] = state;
                if (a.n_priors_out) a.n_priors_out[f] = np;
                __stcg(ictl, np);
                __stcg(ictl + 1, err);
                __threadfence();
            }
            cluster.sync();
            const int np = __ldcg(ictl);
            if (__ldcg(ictl + 1)) st |= ST_TRAVERSE_UB;
            if (!(st & (ST_TOO_FEW_NODES | ST_EMPTY))) {
                // main registration (trackdlo.cpp:998)
                st |= cpd_run<NPASS>(cluster, sm, a, cscr, Xraw, m0, Xc, Yf, Nn, a.sigma2[f], a.sigma2 + f, a.p1,
                                   pri, np, V, nullptr, Nn, a.W ? a.W + (long long)f * Nn * 3 : nullptr,
                                   a.iters ? a.iters + 2 * f + 1 : nullptr);
            }
            if (rank == 0 && tid == 0 && a.status) a.status[f] = st;
        }
    }
}

}  // namespace tdlo
