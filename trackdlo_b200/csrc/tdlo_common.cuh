// Device-side building blocks of the B200-native TrackDLO registration path (sm_100a), shared by the task-queue
// engine (tdlo_taskq.cuh): parameters and frame arguments, the fp64 exp(-z), the prune + counting sort of a chunk of
// points (trackdlo/src/trackdlo.cpp:177-195), the dense solvers that stand in for
// completeOrthogonalDecomposition().solve (trackdlo.cpp:415), the LLE weights (:92-159) and traverse_euclidean
// (:584-898, utils.cpp:172-241).
//
// Everything is fp64 (the reference is MatrixXd end to end; the parity gate is 1e-5 on W).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace tdlo {

constexpr int kMaxNodes = 256;

// status bits (mirror include/trackdlo_b200.h)
constexpr int ST_NOT_CONVERGED = 1, ST_SINGULAR = 2, ST_TOO_FEW_NODES = 4, ST_EMPTY = 8, ST_TRAVERSE_UB = 16,
              ST_PRE_NOT_CONVERGED = 32, ST_BAD_INPUT = 64;

struct CpdP {
    double beta, lambda, gamma, mu, tol, alpha, k_vis, tau, prune_radius;
    int max_iter, include_lle;
};

struct KArgs {
    int mode;            // 0 = batched cpd_lle, 1 = batched tracking_step
    int n_frames;
    int node_stride;     // row stride of Y / W / H (nodes)
    int priors_stride;   // rows per frame of `priors` (mode 0)
    int nmax;            // largest node count in the batch (selects the kernel variant)
    long long max_points;   // capacity of Xc / bkt / the per-chunk arrays: frames whose offsets exceed it are refused
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
    double* packed_out;  // optional [n_frames][3*node_stride + 4]: {Y, sigma2, iters_pre, iters_main, status} per frame (one all-gather)
    CpdP p0;             // mode 0: the call's params; mode 1: pre-processing registration
    CpdP p1;             // mode 1: main registration
    // workspace
    double* Xc;          // compacted points, same indexing as X
    unsigned short* bkt; // nearest-node bucket per raw point (sort key), same indexing as X
    int scr_nodes;       // node capacity the scratch layout was sized for
    unsigned long long* prof;   // optional [16] phase cycle counters
    const double* exp_tab;      // [EXP_TAB] 2^(j/EXP_TAB), filled by the host at context creation
};

// the shared-memory arrays prune_sort_slice works in
struct Smem {
    double4* node4; double* red; double* ptile;
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

// The E-step's exp(-z): 2048-entry table 2^(j/2048) (16 KB of shared memory) + degree-3 polynomial -- two FP64
// instructions fewer than exp_neg; |r| <= ln2/4096 = 1.7e-4, truncation r^4/24 <= 3.4e-17 relative, same error bound and
// clamp as exp_neg.
constexpr int EXP_TAB_BITS = 11, EXP_TAB = 1 << EXP_TAB_BITS;
__device__ __forceinline__ double exp_neg_t(double z, const double* __restrict__ tab) {
    const double L = 2954.639443740597;           // 2048 / ln 2
    const double C_HI = 3.384507717577858e-4;     // ln 2 / 2048
    const double MAGIC = 6755399441055744.0;      // 1.5 * 2^52
    z = __hiloint2double(min(__double2hiint(z), 0x40861000), __double2loint(z));   // NaN also lands here
    const double t = fma(z, -L, MAGIC);
    const int n = __double2loint(t);
    const double nf = t - MAGIC;
    const double r = fma(nf, -C_HI, -z);
    const double tj = tab[n & (EXP_TAB - 1)];
    const double r2 = r * r;
    const double q = fma(r, 1.6666666666666666e-1, 0.5);
    const double p = fma(q, r2, r);
    const double e = fma(tj, p, tj);
    return __hiloint2double(__double2hiint(e) + ((n >> EXP_TAB_BITS) << 20), __double2loint(e));
}

// 32-byte L2 load (data written by other CTAs during this launch must not come from L1)
__device__ __forceinline__ double4 ldcg4(const double4* p) {
    const double2 lo = __ldcg(reinterpret_cast<const double2*>(p)), hi = __ldcg(reinterpret_cast<const double2*>(p) + 1);
    return make_double4(lo.x, lo.y, hi.x, hi.y);
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
static __device__ int prune_sort_slice(const Smem& sm, const double* __restrict__ Xraw, long long r0, long long r1,
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
// Gauss-Jordan elimination with implicit partial (row) pivoting on the augmented system
// AB = [A | B] (n x (n+3), row stride ld).  Replaces completeOrthogonalDecomposition().solve
// (trackdlo.cpp:415) for the full-rank A of this path.  Block-cooperative; W -> wsol[n][3].
// Returns non-zero (uniform) if a zero / non-finite pivot was met.
// ------------------------------------------------------------------------------------------
static __device__ int gj_solve(double* AB, int n, int ld, int* prow, int* used, double* rpiv_slot, double* wsol) {
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
static __device__ int gj_solve_small(double* __restrict__ AB, int n, int ld, double* __restrict__ gj, int* __restrict__ pivs,
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
static __device__ int gj_solve_regs(const double* AB, int n, int ld, double* buf, double* wsol) {   // no __restrict__: buf is written by other threads
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
static __device__ int chol_solve_blocked(double* A, int n, int ld, double* work, double* wsol) {
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
// Structured M-step solve, O(Nn): (diag(D) G + c I) W = B and V = G W without ever forming G.
//
// G_ij = 1/(4 beta^2) exp(-sqrt2 d/beta) (2 d + sqrt2 beta), d = |s_i - s_j| (trackdlo.cpp:225-233), is the Matern-3/2
// covariance sigma_f^2 (1 + a d) exp(-a d) with a = sqrt2/beta, sigma_f^2 = sqrt2/(4 beta), of a process f sampled at the
// (ascending) arc-length coordinates s_i.  (f, f') is a two-dimensional Markov process, so with D = d_i^2 >= 0
//     D^1/2 G D^1/2 + c I  =  Cov(y),   y_i = d_i f(s_i) + noise_i,  noise ~ N(0, c):
// the innovations form of the Kalman filter over the nodes is an L F L^T factorisation of that matrix in O(Nn), its
// adjoint (backward) recursion applies the inverse (de Jong's smoothing error u = Cov(y)^-1 y), and the smoothed mean of
// f is G D^1/2 u.  Hence  W = D^1/2 u  solves (D G + c I) W = B for y = D^-1/2 B  (rows with D_i = 0 have B_i = 0 and get
// W_i = 0, exactly as in the dense SPD form D^1/2 G D^1/2 + cI it replaces), and V = G W comes out of the same backward
// pass -- T = Y0 + G W (trackdlo.cpp:417) needs no dense product either.  Replaces the O(Nn^2) assembly of A
// (trackdlo.cpp:394-413), the O(Nn^3) completeOrthogonalDecomposition().solve (:415) and the O(Nn^2) product G*W.
// Measured against a 50-digit dense solve it is MORE accurate than LAPACK's dense solve of the same system (1e-13 vs
// 1e-12 relative on W at cond 1e5; profiles/r2_kalman_solver_accuracy.txt), because it never forms the ill-conditioned A.
//
// Covariances are carried as the deficit Delta = P_inf - P (starts at 0, only grows), transitions Phi_t over the gap
// h_t = s_{t+1} - s_t are precomputed once per call (phi[t] = {Phi00, Phi01, Phi10, Phi11}).
// All threads first turn D, B into d = sqrt(D) and y = B / d (the square roots stay out of the recursion); then lanes
// 0..2 of warp 0 run the recursion, one right-hand-side column each (the covariance part redundantly: no communication
// inside the ~120-cycle dependent chain of a step; the next node's inputs are loaded one step ahead).
// All pointers are shared memory.
//   in : dd[n] = D_i (overwritten with d_i), bt[3][n] = columns of B (overwritten), phi[n][4], y0[n][3]
//   out: wsol[n][3] = W, tnew[n][3] = Y0 + G W
//   ws : rF[n], kk[2n], pp[2n], am[6n]
// Returns non-zero if an innovation variance was not finite and positive.
// ------------------------------------------------------------------------------------------
static __device__ int mct_kalman_solve(int n, double c, double beta, double* __restrict__ dd, double* __restrict__ bt,
                                       const double* __restrict__ phi, const double* __restrict__ y0, double* __restrict__ wsol,
                                       double* __restrict__ tnew, double* __restrict__ rF, double* __restrict__ kk,
                                       double* __restrict__ pp, double* __restrict__ am) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < n; i += nt) {
        const double D = dd[i];
        const double d = sqrt(D), rd = D > 0.0 ? 1.0 / d : 0.0;
        dd[i] = d;
        bt[i] *= rd; bt[n + i] *= rd; bt[2 * n + i] *= rd;          // y = B / d (0 where D = 0: B is 0 there)
    }
    __syncthreads();
    int bad = 0;
    if (tid < 3) {
        const int col = tid;
        const double a = sqrt(2.0) / beta, s2f = sqrt(2.0) / (4.0 * beta);    // P_inf = diag(s2f, a^2 s2f)
        double* __restrict__ v = bt + col * n;
        double* __restrict__ a0s = am + col * 2 * n;
        double* __restrict__ a1s = a0s + n;
        double D00 = 0.0, D01 = 0.0, D11 = 0.0;          // Delta = P_inf - P (prior of node t before its observation)
        double a0 = 0.0, a1 = 0.0;                        // prior mean of (f, f') at node t
        double d = dd[0], y = v[0];
        double p00 = phi[0], p01 = phi[1], p10 = phi[2], p11 = phi[3];
        for (int t = 0; t < n; t++) {
            // inputs of the next node (independent of the recursion: their latency hides behind this step)
            const int tn = t + 1 < n ? t + 1 : t;
            const double dn = dd[tn], yn = v[tn];
            const double q00 = phi[4 * tn], q01 = phi[4 * tn + 1], q10 = phi[4 * tn + 2], q11 = phi[4 * tn + 3];
            const double P00 = s2f - D00, P01 = -D01;
            const double g0 = P00 * d, g1 = P01 * d;      // P h,  h = (d, 0)
            const double F = fma(d * d, P00, c);          // (d * d does not depend on the recursion)
            const double r = rcp_fast(F);
            bad |= !(F > 0.0) || !(fabs(r) <= 1.79e308);
            const double k0 = g0 * r, k1 = g1 * r;        // gain
            const double inn = fma(-d, a0, y);
            if (col == 0) { rF[t] = r; kk[2 * t] = k0; kk[2 * t + 1] = k1; pp[2 * t] = P00; pp[2 * t + 1] = P01; }
            v[t] = inn; a0s[t] = a0; a1s[t] = a1;
            // posterior after the observation, then transition to node t+1
            const double m0 = fma(k0, inn, a0), m1 = fma(k1, inn, a1);
            const double E00 = fma(g0, k0, D00), E01 = fma(g0, k1, D01), E11 = fma(g1, k1, D11);   // Delta+ = Delta + k k^T F
            a0 = fma(p00, m0, p01 * m1); a1 = fma(p10, m0, p11 * m1);
            const double M00 = fma(p00, E00, p01 * E01), M01 = fma(p00, E01, p01 * E11);
            const double M10 = fma(p10, E00, p11 * E01), M11 = fma(p10, E01, p11 * E11);
            D00 = fma(M00, p00, M01 * p01); D01 = fma(M00, p10, M01 * p11); D11 = fma(M10, p10, M11 * p11);
            d = dn; y = yn; p00 = q00; p01 = q01; p10 = q10; p11 = q11;
        }
        // backward: r <- Phi_t^T r ;  u_t = v_t / F_t - k_t . r ;  r <- r + h u ;  smoothed f_t = a_t[0] + P_t[0,:] . r
        __syncwarp(0x7u);                                 // lanes 1, 2 read what lane 0 stored (rF, kk, pp)
        double r0 = 0.0, r1 = 0.0;
        int t = n - 1;
        double vt = v[t], rf = rF[t], kt0 = kk[2 * t], kt1 = kk[2 * t + 1], pt0 = pp[2 * t], pt1 = pp[2 * t + 1], at = a0s[t], dt = dd[t];
        double yx = y0[3 * t + col];
        for (; t >= 0; t--) {
            const int tp = t > 0 ? t - 1 : 0;
            const double vn = v[tp], rfn = rF[tp], kn0 = kk[2 * tp], kn1 = kk[2 * tp + 1], pn0 = pp[2 * tp], pn1 = pp[2 * tp + 1], an = a0s[tp], dn = dd[tp];
            const double yxn = y0[3 * tp + col];
            const double f00 = phi[4 * tp], f01 = phi[4 * tp + 1], f10 = phi[4 * tp + 2], f11 = phi[4 * tp + 3];    // Phi_{t-1}
            const double u = fma(vt, rf, -fma(kt0, r0, kt1 * r1));
            r0 = fma(dt, u, r0);
            const double V = fma(pt0, r0, fma(pt1, r1, at));
            wsol[3 * t + col] = dt * u;
            tnew[3 * t + col] = yx + V;
            const double n0 = fma(f00, r0, f10 * r1), n1 = fma(f01, r0, f11 * r1);      // adjoint of the transition t-1 -> t
            r0 = n0; r1 = n1;
            vt = vn; rf = rfn; kt0 = kn0; kt1 = kn1; pt0 = pn0; pt1 = pn1; at = an; dt = dn; yx = yxn;
        }
    }
    return __syncthreads_or(bad);
}

// ------------------------------------------------------------------------------------------
// Banded information-form M-step solve WITH the LLE regulariser (pre-processing registration, trackdlo.cpp:396-417), O(Nn).
//
// Let z = (f_0, f'_0, f_1, f'_1, ...) be the joint values of the Matern-3/2 process (whose covariance at the nodes is G) and of
// its derivative, K their precision -- block tridiagonal, because (f, f') is Markov along the arc length -- and P the selection
// of the f components.  With S = diag(D) + eps H  (D = P1 + alpha J, H = E^T E the LLE matrix, eps = sigma2 gamma), c = lambda sigma2:
//     (S G + c I) W = B,   V = G W        <=>        (c K + P^T S P) z = P^T B,    V = P z,    W = P K z.
// The system on the right is symmetric positive definite of size 2 Nn with half-bandwidth 12 (H reaches six nodes): LDL^T
// without pivoting, right-looking on the band with the three right-hand sides carried along, then a back substitution.
// It never forms G, A = S G + cI (condition 1e7..1e10) or a dense factor; measured against a 50-digit dense solve it is
// accurate to 1e-13..1e-16 on W and G W where LAPACK's dense solve of A reaches 1e-9..1e-11
// (scripts/banded_solver_check.py is the NumPy twin of this function, profiles/r2_banded_solver_accuracy.txt).
// Replaces the 50 x 53 pivoted elimination (39 k cycles + the O(Nn^2) assembly) and, above 64 nodes, the 8-state filter.
//
// Per registration (tq_start_call): kh16[2 Nn][16] = lambda K + gamma P^T H P in band storage (slot k of row i = column
// i - 12 + k, the diagonal at k = 12; slots 13..15 zero), kf[Nn][6] = K[2t][2t-2 .. 2t+3].  Per M-step the band is
// sigma2 * kh16 + D on the even diagonal, the right-hand sides sit in slots 13..15 of the even rows.
//   A (shared, 16-byte aligned): [2 Nn][16] + 4 doubles workspace;  dv: D_t of node t = threadIdx.x;  bv[u]: B entry threadIdx.x + u blockDim.x
//   out (shared): wsol[n][3] = W, tnew[n][3] = Y0 + G W.   Returns non-zero if a pivot was not positive and finite.
// ------------------------------------------------------------------------------------------
constexpr int BL_BW = 12, BL_LD = 16;

// exact-to-rounding 1 - exp(-2x)(1 + 2x + 2x^2) (= Q00 / sigma_f^2 of the gap): the closed form cancels to O(x^3)
__device__ inline double bl_q00_factor(double x) {
    if (x >= 0.25) return 1.0 - exp(-2.0 * x) * (1.0 + 2.0 * x + 2.0 * x * x);
    const double y = 2.0 * x;
    double term = y * y * y / 6.0, sum = term;
    for (int k = 4; k < 16; k++) { term *= y / k; sum += term; }
    return exp(-y) * sum;
}

// transition over a gap h: c = Q^-1 Phi (row-major 2x2), ptc = Phi^T Q^-1 Phi (00, 01, 11), qi = Q^-1 (00, 01, 11)
__device__ inline void bl_gap(double h, double a, double s2f, double* c, double* ptc, double* qi) {
    const double x = a * h, e = exp(-x), e2 = e * e;
    const double p00 = e * (1.0 + x), p01 = e * h, p10 = -e * a * x, p11 = e * (1.0 - x);
    const double q00 = s2f * bl_q00_factor(x), q01 = s2f * 2.0 * a * x * x * e2, q11 = s2f * a * a * (1.0 - e2 * (1.0 - 2.0 * x + 2.0 * x * x));
    const double det = q00 * q11 - q01 * q01;
    qi[0] = q11 / det; qi[1] = -q01 / det; qi[2] = q00 / det;
    c[0] = qi[0] * p00 + qi[1] * p10; c[1] = qi[0] * p01 + qi[1] * p11;
    c[2] = qi[1] * p00 + qi[2] * p10; c[3] = qi[1] * p01 + qi[2] * p11;
    ptc[0] = p00 * c[0] + p10 * c[2]; ptc[1] = p00 * c[1] + p10 * c[3]; ptc[2] = p01 * c[1] + p11 * c[3];
}

// Per registration: kh16 and kf (global scratch) from the arc lengths s (shared), H (global, dense [n][n], band |r - c| <= 6).
static __device__ __noinline__ void mct_banded_lle_setup(int n, double beta, double lambda, double gamma, const double* __restrict__ s,
                                            const double* __restrict__ H, double* __restrict__ kh16, double* __restrict__ kf) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const double a = sqrt(2.0) / beta, s2f = sqrt(2.0) / (4.0 * beta);
    for (int t = tid; t < n; t += nt) {
        double* r0 = kh16 + (long long)(2 * t) * BL_LD;
        double* r1 = r0 + BL_LD;
        for (int k = 0; k < 2 * BL_LD; k++) r0[k] = 0.0;
        double cp[4] = {0, 0, 0, 0}, qp[3] = {1.0 / s2f, 0.0, 1.0 / (a * a * s2f)}, tp[3];      // t = 0: the prior P_inf^-1
        double cn[4] = {0, 0, 0, 0}, qn[3], tn[3] = {0, 0, 0};
        if (t > 0) bl_gap(fabs(s[t] - s[t - 1]), a, s2f, cp, tp, qp);
        if (t + 1 < n) bl_gap(fabs(s[t + 1] - s[t]), a, s2f, cn, tn, qn);
        const double k00 = qp[0] + tn[0], k01 = qp[1] + tn[1], k11 = qp[2] + tn[2];          // K_tt
        r0[10] = lambda * -cp[0]; r0[11] = lambda * -cp[1]; r0[12] = lambda * k00;
        r1[9] = lambda * -cp[2]; r1[10] = lambda * -cp[3]; r1[11] = lambda * k01; r1[12] = lambda * k11;
        for (int u = t - 6 > 0 ? t - 6 : 0; u <= t; u++) r0[BL_BW - 2 * (t - u)] += gamma * H[(long long)t * n + u];
        double* f = kf + (long long)t * 6;
        f[0] = -cp[0]; f[1] = -cp[1]; f[2] = k00; f[3] = k01; f[4] = -cn[0]; f[5] = -cn[2];
    }
}

static __device__ __noinline__ int mct_banded_lle_solve(int n, double sigma2, double dv, const double* bv, const double* __restrict__ kh16,
                                           const double* __restrict__ kf, const double* __restrict__ y0, double* __restrict__ A,
                                           double* __restrict__ wsol, double* __restrict__ tnew) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int m = 2 * n;
    // ---- band: sigma2 * kh16 (all loads in flight before the first store)
    {
        constexpr int U = 20;                                  // 16 m / 2 double2 <= 4096 <= 20 * 224
        const double2* __restrict__ src = reinterpret_cast<const double2*>(kh16);
        double2* dst = reinterpret_cast<double2*>(A);
        const int cnt = m * (BL_LD / 2);
        double2 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) { const int i = tid + u * nt; v[u] = i < cnt ? __ldcg(src + i) : make_double2(0.0, 0.0); }
#pragma unroll
        for (int u = 0; u < U; u++) { const int i = tid + u * nt; if (i < cnt) dst[i] = make_double2(sigma2 * v[u].x, sigma2 * v[u].y); }
        for (int i = tid + U * nt; i < cnt; i += nt) { const double2 w = __ldcg(src + i); dst[i] = make_double2(sigma2 * w.x, sigma2 * w.y); }
    }
    __syncthreads();
    if (tid < n) A[(2 * tid) * BL_LD + BL_BW] += dv;
#pragma unroll
    for (int u = 0; u < 3; u++) {
        const int idx = tid + u * nt;
        if (idx < 3 * n) { const int i = idx / 3, d = idx - 3 * i; A[(2 * i) * BL_LD + 13 + d] = bv[u]; }
    }
    __syncthreads();
    int bad = 0;
    // ---- right-looking block LDL^T on the band (2 x 2 pivots = one node (f_t, f'_t) per step: half the dependent chain of a scalar
    // elimination), right-hand sides carried along: warps 0..3, ONE item per thread.  The work of a step is fixed relative to the
    // pivot rows p0 = 2t, p1 = 2t + 1: the 11 rows p0+2 .. p0+12 below them (row p0+13 is a derivative row, whose band ends at
    // distance 3), 66 pairs (i >= j) + 33 right-hand-side entries = 99 items (decoded once).  Per step a thread reads the pivot
    // block, its rows' entries in the two pivot columns (u, v) and its own entry:
    //     entry -= [u_i v_i] P^-1 [u_j v_j]^T,   P^-1 = adj(P) / det(P)   (the adjugate products do not wait for the reciprocal)
    // One named barrier per step: the multipliers [u_i v_i] P^-1 and P^-1 itself overwrite the pivot columns one step LATE, when
    // nobody reads them any more.  (One warp is not an option: an in-order warp issues a dependent instruction every ~7 cycles;
    // the first, four-entries-per-lane version of this loop took 3100 cycles per pivot.)
    if (warp < 4) {
        const int it = tid;
        int ri = 0, rj = 0, cc = -1;
        if (it < 66) {                                         // pair number `it`: rows ri = 2..12, rj = 2..ri
            int r = 1, base = 0;
            while (base + r <= it) { base += r; r++; }
            ri = r + 1; rj = it - base + 2;
        } else if (it < 99) {
            const int w = it - 66;
            ri = w / 3 + 2; cc = w - 3 * (ri - 2);
        }
        const bool item = ri > 0;
        const bool lmul = item && cc < 0 && rj == 2;           // these 11 threads also own the multipliers of row p0 + ri
        // invalid threads read the pivot (harmless) and never write; per-thread pointers step one node (two rows) per pivot
        const double* pi = A + (item ? ri * BL_LD + BL_BW - ri : BL_BW);                       // u_i, v_i = pi[0], pi[1]
        const double* pj0 = A + (item ? (cc < 0 ? rj * BL_LD + BL_BW - rj : 13 + cc) : BL_BW);  // u_j (or the pivot rows' right-hand side)
        const int jstep = (item && cc >= 0) ? BL_LD : 1;                                         // v_j = pj0[jstep]
        double* pc = A + (item ? (cc < 0 ? ri * BL_LD + BL_BW - ri + rj : ri * BL_LD + 13 + cc) : BL_BW);
        double* pd = A + BL_BW;                                 // a = pd[0], b = pd[BL_LD - 1], d = pd[BL_LD]
        double l0p = 0.0, l1p = 0.0, i00p = 0.0, i01p = 0.0, i11p = 0.0;
        bool plive = false;
        for (int t = 0; t < n; t++) {
            const bool live = item && 2 * t + ri < m;
            const double pa = pd[0], pb = pd[BL_LD - 1], pdd = pd[BL_LD];
            const double ui = live ? pi[0] : 0.0, vi = live ? pi[1] : 0.0, uj = live ? pj0[0] : 0.0, vj = live ? pj0[jstep] : 0.0, cur = live ? *pc : 0.0;
            const double det = fma(pa, pdd, -(pb * pb));
            const double q = fma(ui, fma(pdd, uj, -(pb * vj)), vi * fma(pa, vj, -(pb * uj)));
            const double inv = rcp_fast(det);
            if (live) *pc = fma(-q, inv, cur);
            if (t > 0) {                                        // the previous node's columns: its multipliers and its P^-1
                if (lmul && plive) { double* w = const_cast<double*>(pi) - 2 * BL_LD; w[0] = l0p; w[1] = l1p; }
                if (it == 127) { pd[-2 * BL_LD] = i00p; pd[-BL_LD - 1] = i01p; pd[-BL_LD] = i11p; }
            }
            l0p = fma(ui, pdd, -(vi * pb)) * inv; l1p = fma(vi, pa, -(ui * pb)) * inv;
            i00p = pdd * inv; i01p = -pb * inv; i11p = pa * inv; plive = live;
            pd += 2 * BL_LD; pi += 2 * BL_LD; pj0 += 2 * BL_LD; pc += 2 * BL_LD;
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        if (it == 127) { pd[-2 * BL_LD] = i00p; pd[-BL_LD - 1] = i01p; pd[-BL_LD] = i11p; }     // (the last node has no rows below)
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // the diagonal slots now hold the diagonal of every P^-1: positive and finite for a positive definite system
        for (int r = tid; r < m; r += 128) { const double v = A[r * BL_LD + BL_BW]; bad |= !(v > 0.0) || !(v <= 1.79e308); }
        // ---- back substitution z = L^-T D^-1 y, column oriented, one right-hand side per warp (warps 0..2): lane r < 12 owns the
        // pending sum of the columns j = r (mod 12).  Per node: z_t = P^-1 y_t - pending (the two owner lanes), broadcast by
        // shuffles, then every lane adds the two new rows' multipliers times z to its pending column.
        if (warp < 3) {
            const int c = warp;
            double acc = 0.0;
            int o0 = (m - 2) % BL_BW;                          // owner lane of column p0 (p1: o0 + 1)
            for (int t = n - 1; t >= 0; t--) {
                double* __restrict__ A0 = A + (2 * t) * BL_LD;
                const double i00 = A0[BL_BW], i01 = A0[BL_LD + BL_BW - 1], i11 = A0[BL_LD + BL_BW];
                const bool owner = lane == o0 || lane == o0 + 1;       // (only they read y: they overwrite it with z below)
                const double y0v = owner ? A0[13 + c] : 0.0, y1v = owner ? A0[BL_LD + 13 + c] : 0.0;
                int dd = o0 - lane; if (dd <= 0) dd += BL_BW;   // pending column of this lane: p0 - dd (owners: their next column)
                const bool act = lane < BL_BW && 2 * t - dd >= 0;
                const double l0 = act ? A0[BL_BW - dd] : 0.0;                            // row p0, column p0 - dd
                const double l1 = (act && dd < BL_BW) ? A0[BL_LD + BL_BW - 1 - dd] : 0.0;  // row p1, column p0 - dd (distance dd + 1)
                const double zc = (lane == o0 ? fma(i00, y0v, i01 * y1v) : fma(i01, y0v, i11 * y1v)) - acc;     // (meaningful on the owners)
                const double z0 = __shfl_sync(0xffffffffu, zc, o0), z1 = __shfl_sync(0xffffffffu, zc, o0 + 1);
                __syncwarp();
                if (lane == o0) { A0[13 + c] = z0; A0[BL_LD + 13 + c] = z1; }
                if (owner) acc = 0.0;
                acc = fma(l0, z0, fma(l1, z1, acc));
                o0 = o0 == 0 ? BL_BW - 2 : o0 - 2;
            }
        }
    }
    __syncthreads();
    // ---- V = P z, W = P K z
    for (int i = tid; i < 3 * n; i += nt) {
        const int t = i / 3, c = i - 3 * t;
        const double* __restrict__ f = kf + (long long)t * 6;
        double w = 0.0;
#pragma unroll
        for (int q = 0; q < 6; q++) {
            const int j = 2 * t - 2 + q;
            if (j >= 0 && j < m) w = fma(__ldcg(f + q), A[j * BL_LD + 13 + c], w);
        }
        wsol[i] = w;
        tnew[i] = y0[i] + A[(2 * t) * BL_LD + 13 + c];
    }
    return __syncthreads_or(bad);
}

// ------------------------------------------------------------------------------------------
// LLE weights, one node per thread (trackdlo.cpp:92-159).  Mirrors the operation order of
// oracle/trackdlo_oracle.cpp::lle_weights_node with explicitly rounded mul/add/div so that both
// produce identical bits (the 6x6 Gram matrices are rank 3; their inverse is rounding noise).
// Writes row i of E = I - L into E[i*ldE + ...] (row must be pre-zeroed).
// ------------------------------------------------------------------------------------------
static __device__ int lle_lu6(double* a, int nb, int* piv) {
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

static __device__ void lle_row(const double* __restrict__ y0 /*[n][3]*/, int M, int i, double* Erow) {
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

static __device__ bool is_between(V3 x, V3 a, V3 b) {
    const double xs[3] = {x.x, x.y, x.z}, as[3] = {a.x, a.y, a.z}, bs[3] = {b.x, b.y, b.z};
    bool in_bound = true;
    for (int i = 0; i < 3; i++) {
        if (!(as[i] - 0.0001 <= xs[i] && xs[i] <= bs[i] + 0.0001) &&
            !(bs[i] - 0.0001 <= xs[i] && xs[i] <= as[i] + 0.0001)) in_bound = false;
    }
    return in_bound;
}

static __device__ int line_sphere(V3 A, V3 B, V3 C, double radius, V3* out) {
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
static __device__ int pursue(const double* guide, int from, int dir, int bound, V3& centre, double look) {
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
static __device__ int traverse_euclidean(const double* geo, int G, const double* guide, int R, const int* vis, int V,
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

}  // namespace tdlo
