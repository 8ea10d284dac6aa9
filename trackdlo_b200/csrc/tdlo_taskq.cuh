// Task-queue engine of the B200-native TrackDLO registration path (sm_100a).
//
// ONE persistent launch per batch.  Every CTA loops over a global ticket queue of small tasks; a
// frame's EM iteration (trackdlo/src/trackdlo.cpp:276-438) is a wave of independent CHUNK tasks over the
// frame's points followed by the M-step, which the CTA that finishes the wave's last chunk runs on the
// spot ("last arriver continues").  Nothing ever waits on another CTA: while one CTA assembles and solves
// (diag(P1) G + lambda sigma2 I) W = B for a frame, every other CTA keeps streaming E-step chunks of the
// other frames, and the SMs stay busy irrespective of how frames x chunks divide the 148 SMs.
//
//   start_call ──► PRUNE chunk tasks ─┐ (+ the set-up itself: G, LLE, priors)
//                                     └► after_prune ─► begin_iter ─► [DMIN chunk tasks ─► after_dmin] ─►
//                  ESTEP chunk tasks ─► m_step ─► begin_iter ... ─► finish_call ─► (tracking: traverse, main call)
//
// Results are bit-deterministic: chunk partials are combined in chunk order whichever CTA computed them.
// Everything is fp64 (the reference is MatrixXd end to end).
#pragma once

#include "tdlo_common.cuh"

namespace tdlo {

constexpr int TQ_THREADS = 256;            // threads per CTA
constexpr int TQ_ROWS = 32, TQ_RS = 33;    // P tile of a warp: 32 node rows x 32 points (+1 pad)

enum { TK_PRUNE = 1, TK_DMIN = 2, TK_ESTEP = 3, TK_EXIT = 7 };
enum { A_NONE = 0, A_START_CALL, A_AFTER_PRUNE, A_BEGIN_ITER, A_AFTER_DMIN, A_MSTEP, A_FINISH_CALL, A_FRAME_DONE };

// frame control block (ints)
enum { FC_PENDING = 0, FC_PHASE, FC_STAGE, FC_ITER, FC_NN, FC_NCHUNK, FC_USEVIS, FC_NPRI, FC_STATUS, FC_NVIS, FC_STPRE,
       FC_ITPRE, FC_DENSE, FC_WORDS = 16 };
// frame scalars (doubles)
enum { FS_SIGMA2 = 0, FS_RSCALE, FS_CNORM, FS_MP, FS_CGAUSS, FS_WORDS = 8 };

// ---- per-frame scratch (doubles); N = scr_nodes
struct TqScr {
    long long G, HG, H, AB, NODE4, VW, Y0, S, YEXT, JD, HY0, WSOL, SCAL, TRV, PRI, GUIDE, PHI, KTR, KLW, CTL, total;
};
__host__ __device__ inline TqScr tq_scr_layout(int N) {
    TqScr s;
    long long o = 0, n2 = (long long)N * N;
    s.G = o; o += n2;
    s.HG = o; o += n2;
    s.H = o; o += n2;
    s.AB = o; o += (long long)N * (N + 4);
    s.NODE4 = o; o += 4 * N;
    s.VW = o; o += N;
    s.Y0 = o; o += 3 * N;
    s.S = o; o += N;
    s.YEXT = o; o += 3 * N;
    s.JD = o; o += N;
    s.HY0 = o; o += 3 * N;
    s.WSOL = o; o += 3 * N;
    s.SCAL = o; o += FS_WORDS;
    s.TRV = o; o += 2LL * (N + 2) * 4;
    s.PRI = o; o += (2LL * N + 4) * 4;
    s.GUIDE = o; o += 3 * N;
    s.PHI = o; o += 4 * N;
    s.KTR = o; o += 8 * N;                   // structured LLE solve: transitions {Phi, Q} of every gap
    s.KLW = o; o += 43LL * N + 64;            // ... and its forward-pass workspace
    s.CTL = o; o += FC_WORDS / 2;
    s.total = (o + 15) & ~15LL;
    return s;
}

// ---- shared memory layout (bytes).  One fixed head (exp table, node data) + a region that is the E-step's
// P tiles during chunk tasks and the M-step's [A|B] + vectors during continuations.
struct TqSmemL {
    int tab, node4, nsoa, vw, bcast, wbuf, ptile, wacc, y0, s, yext, jd, hy0, p1, px, wsol, tnew, red, gjbuf, prow, used, ab, ab_doubles, chol_doubles, total;
};
__host__ __device__ constexpr TqSmemL tq_smem_layout(int N, int nw) {
    TqSmemL l{};
    int o = 0;
    l.tab = o; o += (N <= 64 ? EXP_TAB : 64) * 8;      // 2048-entry exp table up to 64 nodes, the 64-entry one above (shared memory)
    l.node4 = o; o += N * 32;
    l.nsoa = o; o += N * 8 + 4 * N * 4;  // node s'[] (doubles), then x[], y[], z[] as floats (conservative sphere pruning; lane = node loads)
    l.vw = o; o += N * 8;
    l.bcast = o; o += 64;
    l.red = o; o += 64 * 8;
    o = (o + 31) & ~31;
    const int u = o;
    // E-step view
    l.wbuf = o; o += nw * 32 * 32;
    l.ptile = o; o += nw * TQ_ROWS * TQ_RS * 8;
    // per-warp P1 / PX accumulators [nw][N][4] for N <= 64 (wider node ranges keep them in registers: the P tiles are then
    // reused for the cross-warp reduction)
    l.wacc = o; if (N <= 64) o += nw * N * 32;
    const int e_end = o;
    // M-step view (aliases the E-step view)
    o = u;
    l.y0 = o; o += 3 * N * 8;
    l.s = o; o += N * 8;
    l.p1 = o; o += N * 8;
    l.px = o; o += 3 * N * 8;
    l.wsol = o; o += 3 * N * 8;
    l.tnew = o; o += 3 * N * 8;
    l.gjbuf = o; o += GJR_BUF_DOUBLES * 8;
    l.prow = o; o += N * 4;
    l.used = o; o += N * 4;
    o = (o + 31) & ~31;
    // yext / jd / hy0 are only read while [A|B] is assembled: the blocked Cholesky (Nn > 64) takes its workspace from
    // here to the end of the region
    l.yext = o; o += 3 * N * 8;
    l.jd = o; o += N * 8;
    l.hy0 = o; o += 3 * N * 8;
    o = (o + 31) & ~31;
    l.ab = o;
    l.ab_doubles = (e_end - o) / 8;
    l.chol_doubles = (e_end - l.yext) / 8;
    // the banded LLE solve (mct_banded_lle_solve) takes [2N][16] doubles from gjbuf on (gjbuf / prow / used / yext / jd / hy0 /
    // [A|B] are dead by then)
    const int b_end = l.gjbuf + (2 * N * BL_LD + 8) * 8;
    l.total = e_end > b_end ? e_end : b_end;
    return l;
}

struct TqArgs {
    KArgs k;                         // frame data + parameters
    int chunk;                       // raw points per chunk task
    int inflight;                    // frames started at launch; one more starts whenever a frame completes
    double zcut;                     // Gaussian truncation: entries exp(-z), z > zcut, are skipped (745.2 = exact zeros only)
    double zrel;                     // ... and entries more than exp(-zrel) below the point's largest entry (745.2 = off)
    int solver;                      // M-step solve: 0 = automatic (structured O(Nn) state-space solve without LLE, and with LLE above 64 nodes),
                                     // 1 = dense always, 2 = structured always
    unsigned long long* qctl;        // [0] head ticket, [1] tail, [2] ints {next frame, start tickets}, [3] ints {frames done, abort flag}
    unsigned long long* qslots; unsigned qmask;
    double* fscratch; long long fstride;   // per-frame scratch
    double* part; double* dminp; double* gath; int* nkept;   // per global chunk
    double4* tsph;                   // per tile of 32 sorted points: bounding sphere {centre = first point, radius}; [chunk id][chunk/32]
    int part_stride;                 // doubles per chunk in `part`
    const int* ready; int ready_frames;   // optional: frames [0, *ready * ready_frames) have their points uploaded (pipelined H2D)
    unsigned long long watchdog_ns;  // a CTA that waits longer than this for a task / an upload aborts the launch (0 = never)
    TqSmemL L;
};

struct TqSm {
    double* tab; double4* node4; double* nsoa; double* vw; int* bcast; double* red;
    double4* wbuf; double* ptile; double* wacc;
    double *y0, *s, *yext, *jd, *hy0, *p1, *px, *wsol, *tnew, *gjbuf; int *prow, *used; double* ab;
};

// optional cycle counters (thread 0 of every CTA): 0 queue wait, 1 prune, 2 dmin, 3 E-step, 4 start_call, 5 wave glue
// (after_prune / begin_iter / after_dmin), 6 M-step gather+assemble, 7 solve, 8 update, 9 finish_call (+traversal),
// 10 E-step tasks, 11 tiles, 12 sum of window widths, 13 row blocks
#define TQ_TICK(slot)                                                                                   \
    if (prof && threadIdx.x == 0) { const long long tn_ = clock64(); atomicAdd(prof + (slot), (unsigned long long)(tn_ - tprev)); tprev = tn_; }


// ------------------------------------------------------------------------------------------
// queue
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_acquire_s32(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
// slot word: lap(24) | type(3) | frame(17) | chunk(20)
__device__ __forceinline__ unsigned long long tq_word(unsigned long long ticket, unsigned qmask, int type, int frame, int chunk) {
    const unsigned long long lap = (ticket / ((unsigned long long)qmask + 1ull) + 1ull) & 0xffffffull;
    return (lap << 40) | ((unsigned long long)type << 37) | ((unsigned long long)frame << 20) | (unsigned long long)chunk;
}
// Publishes n tasks (type, frame, chunk 0..n-1).  Called by all threads of the CTA after the data the tasks
// read has been written; contains the fences and a barrier.
static __device__ void tq_push(const TqArgs& a, TqSm& sm, int type, int frame, int n) {
    // release: the CTA barrier orders every thread's writes before thread 0's gpu-scope fence, which is cumulative (the
    // pattern of a cooperative-groups grid barrier) -- one fence instead of one per thread
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); *reinterpret_cast<unsigned long long*>(sm.bcast + 4) = atomicAdd(a.qctl + 1, (unsigned long long)n); }
    __syncthreads();
    const unsigned long long base = *reinterpret_cast<unsigned long long*>(sm.bcast + 4);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned long long t = base + i;
        st_release_u64(a.qslots + (t & a.qmask), tq_word(t, a.qmask, type, type == TK_EXIT ? 0 : frame, type == TK_EXIT ? 0 : i));
    }
    __syncthreads();
}
// Watchdog for the spin loops: a waiter polls it every 1024 spins.  Returns true when the launch is to be abandoned --
// because another CTA said so, or because this wait has exceeded a.watchdog_ns (a task lost, an upload that never
// came): sets the abort flag, which the host turns into TDLO_ERR_CUDA instead of a hung caller.
__device__ __forceinline__ unsigned long long tq_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ bool tq_watchdog(const TqArgs& a, unsigned long long& t0) {
    int* abort_flag = reinterpret_cast<int*>(a.qctl + 3) + 1;
    if (ld_acquire_s32(abort_flag)) return true;
    if (!a.watchdog_ns) return false;
    const unsigned long long now = tq_now_ns();
    if (!t0) { t0 = now; return false; }
    if (now - t0 > a.watchdog_ns) { atomicExch(abort_flag, 1); return true; }
    return false;
}
// Takes the next ticket and waits for its task.  Returns the slot word (uniform over the CTA).
static __device__ unsigned long long tq_pop(const TqArgs& a, TqSm& sm) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long t = atomicAdd(a.qctl, 1ull);
        const unsigned long long lap = (t / ((unsigned long long)a.qmask + 1ull) + 1ull) & 0xffffffull;
        const unsigned long long* slot = a.qslots + (t & a.qmask);
        unsigned long long v = ld_acquire_u64(slot);
        unsigned ns = 64, spins = 0;
        unsigned long long t0 = 0;
        while ((v >> 40) != lap) {
            __nanosleep(ns); if (ns < 512) ns <<= 1;
            v = ld_acquire_u64(slot);
            if ((++spins & 1023u) == 0 && tq_watchdog(a, t0)) { v = (lap << 40) | ((unsigned long long)TK_EXIT << 37); break; }
        }
        *reinterpret_cast<unsigned long long*>(sm.bcast + 4) = v;
    }
    __syncthreads();
    const unsigned long long v = *reinterpret_cast<unsigned long long*>(sm.bcast + 4);
    __syncthreads();
    return v;
}
// A chunk task (or the set-up) of frame f is complete.  Returns true (uniform) for the LAST arriver of the
// wave, which then owns the frame until it publishes the next wave.
static __device__ bool tq_arrive(TqSm& sm, int* ctl) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();                                      // release (cumulative over the CTA's writes, see tq_push)
        const int old = atomicSub(ctl + FC_PENDING, 1);
        sm.bcast[0] = (old == 1);
        if (old == 1) __threadfence();
    }
    __syncthreads();
    const bool last = sm.bcast[0] != 0;
    __syncthreads();
    return last;
}

__device__ __forceinline__ int tq_chunk_base(const TqArgs& a, int f) { return (int)(a.k.x_off[f] / a.chunk) + f; }

// positive doubles order like their bit patterns: warp min / max of the HIGH words gives a bound that is
// conservative by < 2^-20 relative -- good enough for search ranges and windows, one REDUX instead of a
// ten-shuffle tree.  (Both return a value <= / >= the true extremum.)
__device__ __forceinline__ double warp_min_pos_lb(double v) {
    const unsigned h = __reduce_min_sync(0xffffffffu, (unsigned)__double2hiint(v));
    return __hiloint2double((int)h, 0);
}
__device__ __forceinline__ double warp_max_pos_ub(double v) {
    const unsigned h = __reduce_max_sync(0xffffffffu, (unsigned)__double2hiint(v));
    return __hiloint2double((int)h + 1, 0);
}
__device__ __forceinline__ double warp_min_pos_ub(double v) {
    const unsigned h = __reduce_min_sync(0xffffffffu, (unsigned)__double2hiint(v));
    return __hiloint2double((int)h + 1, 0);
}

// Bounding spheres of the chunk's tiles (32 consecutive sorted points): centre = the tile's first point,
// radius = an upper bound of the largest distance to it.  The points do not move during a registration,
// so this is done once per call, right after the prune/sort; the E-step uses it to prune its node searches.
static __device__ void tq_tile_spheres(const double* __restrict__ Xc, int n_local, double4* __restrict__ sph) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int base = warp * 32; base < n_local; base += nw * 32) {
        const int n = base + lane < n_local ? base + lane : base;
        const double x = Xc[(long long)n * 3], y = Xc[(long long)n * 3 + 1], z = Xc[(long long)n * 3 + 2];   // written by this CTA
        const double cx = __shfl_sync(0xffffffffu, x, 0), cy = __shfl_sync(0xffffffffu, y, 0), cz = __shfl_sync(0xffffffffu, z, 0);
        const double rho = sqrt(warp_max_pos_ub(dist2(x, y, z, cx, cy, cz) + 1e-300)) * (1.0 + 1e-9);
        if (lane == 0) sph[base >> 5] = make_double4(cx, cy, cz, rho);
    }
}

// Upper bound of sqrt(x), x > 0 (three instructions; only used for conservative search bounds).
__device__ __forceinline__ double sqrt_ub(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return x * r * (1.0 + 1e-5);
}

// exp(-z) of the E-step: the 2048-entry table variant where the shared-memory budget allows it (node ranges up to 64)
template <int NPASS>
__device__ __forceinline__ double tq_exp(double z, const double* __restrict__ tab) { return NPASS <= 2 ? exp_neg_t(z, tab) : exp_neg(z, tab); }

// Phase B of the E-step for one block of Wb <= WP node rows (WP = 8, 16, 32): lane = (row r, point group g), every lane
// accumulates P1 / PX of its row over the WP points of its group; the 32 / WP groups are then combined by shuffles, so
// that every lane ends up with the sums of row r over all 32 points.
template <int WP>
__device__ __forceinline__ void tq_phase_b(const double* __restrict__ pt, const double2* __restrict__ wlo, const double2* __restrict__ whi,
                                           int lane, int Wb, double& b0, double& b1, double& b2, double& b3) {
    const int r = lane & (WP - 1);
    const double* __restrict__ prow = pt + r * TQ_RS + (lane - r);       // points [g*WP, (g+1)*WP)
    const double2* __restrict__ wgl = wlo + (lane - r);
    const double2* __restrict__ wgh = whi + (lane - r);
    b0 = 0.0; b1 = 0.0; b2 = 0.0; b3 = 0.0;
    if (r < Wb) {
        // two interleaved partial sums per accumulator (even / odd points): eight independent DFMA chains -- four do not cover
        // the FP64 latency (this line carried 9 % of the kernel's stall samples)
        double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
#pragma unroll
        for (int u = 0; u < WP; u += 2) {
            const double p = prow[u], q = prow[u + 1];
            const double2 wa = wgl[u], wc = wgh[u], xa = wgl[u + 1], xc = wgh[u + 1];
            b0 = fma(p, wa.x, b0); b1 = fma(p, wa.y, b1); b2 = fma(p, wc.x, b2); b3 = fma(p, wc.y, b3);
            c0 = fma(q, xa.x, c0); c1 = fma(q, xa.y, c1); c2 = fma(q, xc.x, c2); c3 = fma(q, xc.y, c3);
        }
        b0 += c0; b1 += c1; b2 += c2; b3 += c3;
    }
#pragma unroll
    for (int off = WP; off < 32; off <<= 1) {
        b0 += __shfl_xor_sync(0xffffffffu, b0, off); b1 += __shfl_xor_sync(0xffffffffu, b1, off);
        b2 += __shfl_xor_sync(0xffffffffu, b2, off); b3 += __shfl_xor_sync(0xffffffffu, b3, off);
    }
}

// ------------------------------------------------------------------------------------------
// Fused E-step over one chunk (trackdlo.cpp:278-389): distances -> arg-max node -> geodesic distances -> P ->
// (visibility weights) -> normalisation -> P1, PX, sum Pt1*|x|^2.  Every WARP is autonomous: it takes 32 sorted
// points at a time (lane = point), writes their P columns into its private shared-memory tile (phase A), then
// accumulates P1 / PX of the tile's node rows in registers (phase B).  Only __syncwarp inside the tile loop.
// (1) The P tile of a warp has a fixed 32 node rows -- the node WINDOW of the warp's 32 points (everything outside
// is exactly 0 / below the truncation) is processed in blocks of 32 rows, recomputing the exponentials of later
// blocks (only the first, wide iterations of a registration need more than one); (2) phase B splits the lanes into
// (node row, point group) so that a narrow window still uses all 32 lanes; (3) window and search bounds use
// single-REDUX conservative bounds.
// part_out: [Nn][4] = {P1, PX.x, PX.y, PX.z}, then [4*Nn] = sum_n Pt1_n |x_n|^2.
// ------------------------------------------------------------------------------------------
template <int NPASS, bool VIS, int NW>
static __device__ void tq_estep_chunk(const TqSm& sm, const double* __restrict__ Xc, const double4* __restrict__ sph, int n_local, int Nn,
                               double sigma2, double c_norm, double rscale, double zcut, double zrel, double k_vis, double* part_out,
                               unsigned long long* prof) {
    constexpr int RS = TQ_RS;
    const int tid = threadIdx.x;
    int lane = tid & 31, warp = tid >> 5;
    // opaque to the compiler: otherwise every use inside the tile loop re-derives them from %tid (S2R + LOP3 + IMAD chains,
    // ~60 instructions per tile in the ncu source view) instead of keeping two registers
    asm volatile("" : "+r"(lane), "+r"(warp));
    constexpr int nw = NW, nt = NW * 32;
    double* __restrict__ pt = sm.ptile + warp * (TQ_ROWS * RS);
    double2* __restrict__ wlo = reinterpret_cast<double2*>(sm.wbuf + warp * 32);     // (w, w x) per point
    double2* __restrict__ whi = wlo + 32;                                               // (w y, w z) per point
    double* __restrict__ pcol = pt + lane;
    const double* __restrict__ tab = sm.tab;
    const double4* __restrict__ nd = sm.node4;
    const double* __restrict__ vw = sm.vw;
    const double* __restrict__ nss = sm.nsoa;
    const float* __restrict__ nfx = reinterpret_cast<const float*>(sm.nsoa + 32 * NPASS);
    const float* __restrict__ nfy = nfx + 32 * NPASS;
    const float* __restrict__ nfz = nfy + 32 * NPASS;
    // P1 / PX accumulators of this warp: shared memory [Nn][4] for Nn <= 64 (frees 32 registers in the tile loop and the
    // hand-over to owner lanes), registers (lane l owns nodes l, l+32, ...) above
    constexpr bool SACC = NPASS <= 2;
    double4* __restrict__ wacc = reinterpret_cast<double4*>(sm.wacc) + warp * (32 * NPASS);
    double acc[SACC ? 1 : NPASS][4];
    if (SACC) {
        for (int m = lane; m < Nn; m += 32) wacc[m] = make_double4(0.0, 0.0, 0.0, 0.0);
        __syncwarp();
    } else {
#pragma unroll
        for (int ps = 0; ps < (SACC ? 1 : NPASS); ps++) { acc[ps][0] = acc[ps][1] = acc[ps][2] = acc[ps][3] = 0.0; }
    }
    double sxx = 0.0;
    const double uflow = 1490.2 * sigma2;
    const double T = sqrt(zcut);
    const double s_last = nd[Nn - 1].w;

    const long long t_loop0 = prof ? clock64() : 0;
    // software prefetch: the next tile's point and sphere are requested before the current tile is processed
    double xn = 0.0, yn = 0.0, zn = 0.0;
    double4 sn = make_double4(0.0, 0.0, 0.0, 0.0);
    const double* __restrict__ xp = Xc + (warp * 32 + lane) * 3;         // this lane's point of the warp's next tile
    const double4* __restrict__ sp = sph + warp;
    if (warp * 32 < n_local) {
        const double* q = warp * 32 + lane < n_local ? xp : xp - lane * 3;
        xn = __ldcg(q); yn = __ldcg(q + 1); zn = __ldcg(q + 2);
        sn = ldcg4(sp);
    }
    for (int base = warp * 32; base < n_local; base += nw * 32) {
        const bool valid = base + lane < n_local;            // idle lanes shadow the tile's first point (weight 0)
        const double x = xn, y = yn, z = zn;
        const double cx = sn.x, cy = sn.y, cz = sn.z, rho = sn.w;
        xp += nw * 96; sp += nw;
        if (base + nw * 32 < n_local) {
            const double* q = base + nw * 32 + lane < n_local ? xp : xp - lane * 3;
            xn = __ldcg(q); yn = __ldcg(q + 1); zn = __ldcg(q + 2);
            sn = ldcg4(sp);
        }

        // ---- nearest node: bounding-sphere pruning of the scan range: a node farther from the tile's sphere {centre c,
        // radius rho} (from the prune pass) than the nearest centre distance + 2 rho cannot be nearest to any of the 32
        // points.  The range only has to be CONSERVATIVE (the scan below is exact fp64), so this runs in fp32 with the
        // rounding of the inputs (<= 3e-7 |c|_1 absolute) and of the arithmetic (<= 1e-5 relative) added to the limit.
        int ja, jb;
        {
            const float cxf = (float)cx, cyf = (float)cy, czf = (float)cz;
            const float slack = 3e-7f * (fabsf(cxf) + fabsf(cyf) + fabsf(czf));
            float dc[NPASS];
            float dloc = 3.0e38f;
#pragma unroll
            for (int ps = 0; ps < NPASS; ps++) {
                const int m = lane + 32 * ps;
                dc[ps] = 3.0e38f;
                if (m < Nn) { const float ex = nfx[m] - cxf, ey = nfy[m] - cyf, ez = nfz[m] - czf; dc[ps] = ex * ex + ey * ey + ez * ez; }
                dloc = fminf(dloc, dc[ps]);
            }
            const float dnear = __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(dloc)));     // non-negative floats order like their bits
            float lim = (sqrtf(dnear) * (1.0f + 1e-5f) + 2.0f * ((float)rho * (1.0f + 1e-6f)) + 3.0f * slack) * (1.0f + 1e-5f);
            lim = lim * lim * (1.0f + 1e-6f);
            ja = Nn; jb = -1;
#pragma unroll
            for (int ps = 0; ps < NPASS; ps++) {
                const unsigned mk = __ballot_sync(0xffffffffu, lane + 32 * ps < Nn && dc[ps] <= lim);
                if (mk) { if (ja == Nn) ja = 32 * ps + __ffs(mk) - 1; jb = 32 * ps + 31 - __clz(mk); }
            }
        }
        double best = 1e300;
        int a = ja;
        {
            int j = ja;
            for (; j + 1 <= jb; j += 2) {
                const double4 q0 = nd[j], q1 = nd[j + 1];
                const double e0 = dist2(q0.x, q0.y, q0.z, x, y, z), e1 = dist2(q1.x, q1.y, q1.z, x, y, z);
                if (e0 < best) { best = e0; a = j; }
                if (e1 < best) { best = e1; a = j + 1; }
            }
            if (j <= jb) {
                const double4 q = nd[j];
                const double d2 = dist2(q.x, q.y, q.z, x, y, z);
                if (d2 < best) { best = d2; a = j; }
            }
        }
        // whole column underflows to 0 in the reference -> maxCoeff returns index 0 (trackdlo.cpp:310)
        if (best > uflow && (-0.5 * best) / sigma2 < -745.1332191019412) {
            a = 0;
            const double4 q = nd[0];
            best = dist2(q.x, q.y, q.z, x, y, z);
        }
        a = min(a, Nn - 1);                              // (only reachable with non-finite coordinates)
        int q1 = a - 1; if (q1 == -1) q1 = 2;
        int q2 = a + 1; if (q2 == Nn) q2 = Nn - 3;
        double e1, e2;
        { const double4 q = nd[q1]; e1 = dist2(q.x, q.y, q.z, x, y, z); }
        { const double4 q = nd[q2]; e2 = dist2(q.x, q.y, q.z, x, y, z); }
        // trackdlo.cpp:324-329 compares the two Euclidean distances; comparing their squares picks the same
        // neighbour unless the two distances agree to the last bit
        const bool pick1 = e1 < e2;
        const int b = pick1 ? q1 : q2;
        const double da = sqrt(best);
        const double db = sqrt(pick1 ? e1 : e2);
        const int lo = a < b ? a : b, hi = a < b ? b : a;
        const double dlo = a < b ? da : db, dhi = a < b ? db : da;
        const double alo = nd[lo].w + dlo * rscale;      // t_j = alo - s'_j  for j <= lo
        const double ahi = dhi * rscale - nd[hi].w;      // t_j = ahi + s'_j  for j >= hi

        // ---- node window [jlo, jhi] of this warp: P entries outside are, for each of the 32 points, below exp(-zcut) or more
        // than exp(-zrel) below the point's largest entry (>= exp(-zn) vw_a, a = its nearest node) -- with the default
        // zrel = 45 that is 3e-20 of the column sum they would enter, three orders below half an ulp.  With visibility
        // weights: vw_j / vw_a <= exp(k_vis dm_a) <= exp(k_vis da) (node a is at most da away from a point), added to the margin.
        int jlo, jhi;
        {
            double zn = fma(best, rscale * rscale, zrel);
            if (VIS) zn = fma(fabs(k_vis), da, zn);
            const double tl = fmin(sqrt_ub(zn), T);                                 // this point keeps |t| <= tl
            const double thr_lo = warp_min_pos_lb(alo + (T - tl)) - T;                        // keep j <= lo while s'_j > thr_lo
            const double thr_hi = (T + s_last) - warp_min_pos_lb((ahi + s_last) + (T - tl));  // keep j >= hi while s'_j < thr_hi
            // s' is non-decreasing in j: the window ends are found by looking 32 rows at a time below the smallest lo /
            // above the largest hi of the tile (one ballot each in all but the first, wide iterations)
            jlo = __reduce_min_sync(0xffffffffu, lo);
            jhi = __reduce_max_sync(0xffffffffu, hi);
            for (;;) {
                const int j = jlo - 1 - lane;
                const unsigned m = __ballot_sync(0xffffffffu, j >= 0 && nss[max(j, 0)] > thr_lo);
                const int k = __ffs(~m) - 1;                 // leading lanes that keep their row (-1: all 32)
                if (k >= 0) { jlo -= k; break; }
                jlo -= 32;
            }
            for (;;) {
                const int j = jhi + 1 + lane;
                const unsigned m = __ballot_sync(0xffffffffu, j < Nn && nss[min(j, Nn - 1)] < thr_hi);
                const int k = __ffs(~m) - 1;
                if (k >= 0) { jhi += k; break; }
                jhi += 32;
            }
        }
        if (prof && lane == 0) { atomicAdd(prof + 11, 1ull); atomicAdd(prof + 12, (unsigned long long)(jhi - jlo + 1)); atomicAdd(prof + 13, (unsigned long long)((jhi - jlo) / TQ_ROWS + 1)); }
        const double nahi = -ahi;
        const bool quirk = (hi - lo == 2);               // the row strictly between lo and hi keeps geodesic 0 -> P = 1 (x vw)
        const int jq = lo + 1;

        double w = 0.0;
        for (int j0 = jlo; j0 <= jhi; j0 += TQ_ROWS) {
            const int j1 = min(j0 + TQ_ROWS - 1, jhi);
            // ---- phase A (trackdlo.cpp:332-354, 358-375): P column entries of rows [j0, j1] -> tile; the first
            // block also runs over the rest of the window to complete the column sum.
            const int jend = (j0 == jlo) ? jhi : j1;
            double colsum = 0.0;
            {
                double* pc = pcol;
                int j = j0;
                for (; j + 3 <= jend; j += 4) {
                    const double s0 = nd[j].w, s1 = nd[j + 1].w, s2 = nd[j + 2].w, s3 = nd[j + 3].w;
                    double v0 = 1.0, v1 = 1.0, v2 = 1.0, v3 = 1.0;
                    if (VIS) { v0 = vw[j]; v1 = vw[j + 1]; v2 = vw[j + 2]; v3 = vw[j + 3]; }
                    // t_j = alo - s'_j (j <= lo) or ahi + s'_j (j > lo); only t^2 is used: select the centre, not the result
                    const double t0 = ((j <= lo) ? alo : nahi) - s0;
                    const double t1 = ((j + 1 <= lo) ? alo : nahi) - s1;
                    const double t2 = ((j + 2 <= lo) ? alo : nahi) - s2;
                    const double t3 = ((j + 3 <= lo) ? alo : nahi) - s3;
                    double p0 = tq_exp<NPASS>(t0 * t0, tab), p1 = tq_exp<NPASS>(t1 * t1, tab), p2 = tq_exp<NPASS>(t2 * t2, tab), p3 = tq_exp<NPASS>(t3 * t3, tab);
                    if (VIS) { p0 *= v0; p1 *= v1; p2 *= v2; p3 *= v3; }
                    colsum += (p0 + p1) + (p2 + p3);
                    if (j + 3 <= j1) { pc[0] = p0; pc[RS] = p1; pc[2 * RS] = p2; pc[3 * RS] = p3; }
                    else {
                        if (j <= j1) pc[0] = p0;
                        if (j + 1 <= j1) pc[RS] = p1;
                        if (j + 2 <= j1) pc[2 * RS] = p2;
                    }
                    pc += 4 * RS;
                }
                for (; j <= jend; j++) {
                    const double sj = nd[j].w;
                    const double t = ((j <= lo) ? alo : nahi) - sj;
                    double p = tq_exp<NPASS>(t * t, tab);
                    if (VIS) p *= vw[j];
                    colsum += p;
                    if (j <= j1) *pc = p;
                    pc += RS;
                }
            }
            if (quirk) {
                const double pn = VIS ? vw[jq] : 1.0;
                if (j0 == jlo) {
                    const double tq = ahi + nd[jq].w;    // what the loop computed for row jq (jq > lo)
                    double pq = tq_exp<NPASS>(tq * tq, tab);
                    if (VIS) pq *= vw[jq];
                    colsum += pn - pq;
                }
                if (jq >= j0 && jq <= j1) pcol[(jq - j0) * RS] = pn;
            }
            if (j0 == jlo) {
                const double den = colsum + c_norm;      // trackdlo.cpp:379 / 382
                w = valid ? 1.0 / den : 0.0;
                sxx = fma(colsum * w, x * x + y * y + z * z, sxx);   // Pt1_n * |x_n|^2 (trackdlo.cpp:418)
                wlo[lane] = make_double2(w, w * x); whi[lane] = make_double2(w * y, w * z);
            }
            __syncwarp();

            // ---- phase B (trackdlo.cpp:387-389): lane = (row r, point group g); P1 / PX of the block's rows over
            // the warp's 32 points; groups are combined by shuffles and handed to the lanes that own the nodes.
            const int Wb = j1 - j0 + 1;
            double b0, b1, b2, b3;
            if (Wb <= 8) tq_phase_b<8>(pt, wlo, whi, lane, Wb, b0, b1, b2, b3);
            else if (Wb <= 16) tq_phase_b<16>(pt, wlo, whi, lane, Wb, b0, b1, b2, b3);
            else tq_phase_b<32>(pt, wlo, whi, lane, Wb, b0, b1, b2, b3);
            if (SACC) {
                // lanes 0 .. Wb-1 (row r = lane, group 0) hold the block's row sums
                if (lane < Wb) {
                    double4 v = wacc[j0 + lane];
                    v.x += b0; v.y += b1; v.z += b2; v.w += b3;
                    wacc[j0 + lane] = v;
                }
            } else {
                // owner lane l holds nodes l, l+32, ...; at most one of them lies in [j0, j1]
                const int src = (lane - j0) & 31;             // row index of the owned node inside this block, if any
                const double o0 = __shfl_sync(0xffffffffu, b0, src), o1 = __shfl_sync(0xffffffffu, b1, src);
                const double o2 = __shfl_sync(0xffffffffu, b2, src), o3 = __shfl_sync(0xffffffffu, b3, src);
                const int m = j0 + src;                       // the owned node (m % 32 == lane)
                if (src < Wb) {
#pragma unroll
                    for (int ps = 0; ps < (SACC ? 1 : NPASS); ps++)
                        if ((m >> 5) == ps) { acc[ps][0] += o0; acc[ps][1] += o1; acc[ps][2] += o2; acc[ps][3] += o3; }
                }
            }
            __syncwarp();
        }
    }

    if (prof && lane == 0) atomicAdd(prof + 14, (unsigned long long)(clock64() - t_loop0));     // per-warp tile-loop cycles
    const long long t_red0 = prof ? clock64() : 0;
    // ---- cross-warp reduction in a fixed order (deterministic)
    __syncthreads();
    if (SACC) {
        const double* __restrict__ racc = sm.wacc;        // [nw][32 NPASS][4]
        for (int i = tid; i < 4 * Nn; i += nt) {
            double v = 0.0;
#pragma unroll
            for (int ww = 0; ww < nw; ww++) v += racc[ww * (128 * NPASS) + i];
            __stcg(part_out + i, v);
        }
    } else {
        double* __restrict__ racc = sm.ptile;             // [nw][Nn][4]; the P tiles are dead now
#pragma unroll
        for (int ps = 0; ps < (SACC ? 1 : NPASS); ps++) {
            const int m = lane + 32 * ps;
            if (m < Nn) {
                double* dst = racc + ((long long)warp * Nn + m) * 4;
                dst[0] = acc[ps][0]; dst[1] = acc[ps][1]; dst[2] = acc[ps][2]; dst[3] = acc[ps][3];
            }
        }
        __syncthreads();
        for (int i = tid; i < 4 * Nn; i += nt) {
            double v = 0.0;
            for (int ww = 0; ww < nw; ww++) v += racc[ww * 4 * Nn + i];
            __stcg(part_out + i, v);
        }
    }
    const double sx = block_sum(sxx, sm.red);
    if (tid == 0) __stcg(part_out + 4 * Nn, sx);
    __syncthreads();
    if (prof && tid == 0) atomicAdd(prof + 15, (unsigned long long)(clock64() - t_red0));         // barrier wait + reduction (thread 0)
}

// ------------------------------------------------------------------------------------------
// Visibility pre-pass over one chunk: per-node min squared distance to the chunk's points
// (trackdlo.cpp:279-296).  Exact, with bounding-sphere pruning: a node whose distance to the tile's
// sphere already exceeds the best distance found so far cannot improve it.
// ------------------------------------------------------------------------------------------
template <int NPASS>
static __device__ void tq_dmin_chunk(const TqSm& sm, const double* __restrict__ Xc, int n_local, int Nn, double* dmin_out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5, nt = blockDim.x;
    const double4* __restrict__ nd = sm.node4;
    double best[NPASS];                                   // lane owns nodes lane + 32 ps
#pragma unroll
    for (int ps = 0; ps < NPASS; ps++) best[ps] = 1e300;
    // seed: distance to the first point of every tile of this warp (cheap upper bounds)
    for (int base = warp * 32; base < n_local; base += nw * 32) {
        const double cx = __ldcg(Xc + (long long)base * 3), cy = __ldcg(Xc + (long long)base * 3 + 1), cz = __ldcg(Xc + (long long)base * 3 + 2);
#pragma unroll
        for (int ps = 0; ps < NPASS; ps++) {
            const int m = lane + 32 * ps;
            if (m < Nn) { const double4 q = nd[m]; best[ps] = fmin(best[ps], dist2(q.x, q.y, q.z, cx, cy, cz)); }
        }
    }
    for (int base = warp * 32; base < n_local; base += nw * 32) {
        const bool valid = base + lane < n_local;
        const int n = valid ? base + lane : base;
        const double x = __ldcg(Xc + (long long)n * 3), y = __ldcg(Xc + (long long)n * 3 + 1), z = __ldcg(Xc + (long long)n * 3 + 2);
        const double cx = __shfl_sync(0xffffffffu, x, 0), cy = __shfl_sync(0xffffffffu, y, 0), cz = __shfl_sync(0xffffffffu, z, 0);
        const double rho = sqrt(warp_max_pos_ub(dist2(x, y, z, cx, cy, cz) + 1e-300)) * (1.0 + 1e-9);
#pragma unroll
        for (int ps = 0; ps < NPASS; ps++) {
            const int m = lane + 32 * ps;
            bool cand = false;
            if (m < Nn) {
                const double4 q = nd[m];
                const double dcn = sqrt(dist2(q.x, q.y, q.z, cx, cy, cz));
                const double lb = dcn - rho;                                 // every point of the tile is at least this far
                cand = !(lb > 0.0 && lb * lb * (1.0 - 1e-9) > best[ps]);
            }
            unsigned mk = __ballot_sync(0xffffffffu, cand);
            while (mk) {
                const int l = __ffs(mk) - 1;
                mk &= mk - 1;
                const double4 q = nd[l + 32 * ps];
                const double d2 = dist2(q.x, q.y, q.z, x, y, z);            // idle lanes shadow point 0: harmless for a min
                const unsigned hi = (unsigned)__double2hiint(d2);
                const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
                const unsigned lw = (hi == mh) ? (unsigned)__double2loint(d2) : 0xffffffffu;
                const unsigned ml = __reduce_min_sync(0xffffffffu, lw);
                if (lane == l) best[ps] = fmin(best[ps], __hiloint2double((int)mh, (int)ml));
            }
        }
    }
    __syncthreads();
    double* wmin = sm.ptile;                              // [nw][Nn]
#pragma unroll
    for (int ps = 0; ps < NPASS; ps++) { const int m = lane + 32 * ps; if (m < Nn) wmin[warp * Nn + m] = best[ps]; }
    __syncthreads();
    for (int j = tid; j < Nn; j += nt) {
        double m = wmin[j];
        for (int w = 1; w < nw; w++) m = fmin(m, wmin[w * Nn + j]);
        __stcg(dmin_out + j, m);
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// frame helpers
// ------------------------------------------------------------------------------------------
struct TqFrame {
    int f;
    double* scr; TqScr sc; int* ctl; double* scal;
    int gbase;               // first global chunk id of the frame
    const double* Xraw; double* Xc; long long m0;
};
__device__ __forceinline__ TqFrame tq_frame(const TqArgs& a, int f) {
    TqFrame fr;
    fr.f = f;
    fr.sc = tq_scr_layout(a.k.scr_nodes);
    fr.scr = a.fscratch + (long long)f * a.fstride;
    fr.ctl = reinterpret_cast<int*>(fr.scr + fr.sc.CTL);
    fr.scal = fr.scr + fr.sc.SCAL;
    fr.gbase = tq_chunk_base(a, f);
    const long long x0 = a.k.x_off[f];
    fr.m0 = a.k.x_off[f + 1] - x0;
    fr.Xraw = a.k.X + x0 * 3;
    fr.Xc = a.k.Xc + x0 * 3;
    return fr;
}
__device__ __forceinline__ const CpdP& tq_params(const TqArgs& a, int stage) { return (a.k.mode == 1 && stage == 1) ? a.k.p1 : a.k.p0; }

// global in/out node array of the current call
__device__ __forceinline__ double* tq_yio(const TqArgs& a, const TqFrame& fr, int stage) {
    if (a.k.mode == 0) return a.k.Y + (long long)fr.f * a.k.node_stride * 3;
    if (stage == 0) return a.k.guide_out ? a.k.guide_out + (long long)fr.f * a.k.node_stride * 3 : fr.scr + fr.sc.GUIDE;
    return a.k.Y + (long long)fr.f * a.k.node_stride * 3;
}
__device__ __forceinline__ double* tq_pri(const TqArgs& a, const TqFrame& fr) {
    return a.k.priors_out ? a.k.priors_out + (long long)fr.f * 2 * a.k.node_stride * 4 : fr.scr + fr.sc.PRI;
}

// ------------------------------------------------------------------------------------------
// start_call: set-up of one cpd_lle call (trackdlo.cpp:197-260) + publication of its PRUNE wave.
// Returns the next action for this CTA.
// ------------------------------------------------------------------------------------------
static __device__ int tq_start_call(const TqArgs& a, TqSm& sm, const TqFrame& fr, int stage, unsigned long long& local_wd) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const KArgs& k = a.k;
    const CpdP& p = tq_params(a, stage);
    const int f = fr.f;
    double* scr = fr.scr;
    const TqScr& sc = fr.sc;
    if (a.ready && stage == 0) {           // points of this frame still in flight on the copy stream? (host-buffer entry points)
        if (tid == 0) {
            unsigned ns = 256, spins = 0;
            unsigned long long t0 = 0;
            while ((long long)ld_acquire_s32(a.ready) * a.ready_frames <= f) {
                __nanosleep(ns); if (ns < 4096) ns <<= 1;
                if ((++spins & 1023u) == 0 && tq_watchdog(a, t0)) break;
            }
        }
        __syncthreads();
    }
    // ---- the caller's arrays are device memory nobody has looked at yet (device-pointer entry points): refuse a frame
    // whose offsets / counts / indices would make this kernel write outside the context's workspace
    if (stage == 0) {
        if (tid == 0) {
            bool bad = false;
            const long long x0 = k.x_off[f], x1 = k.x_off[f + 1];
            if (x0 < 0 || x1 < x0 || x1 > k.max_points) bad = true;
            if (k.mode == 0) {
                const int nn = k.n_nodes ? k.n_nodes[f] : k.node_stride;
                if (nn < 0 || nn > k.node_stride) bad = true;
            } else {
                const int N = k.node_stride;
                const long long V = k.ext_off[f + 1] - k.ext_off[f], nv = k.vis_off[f + 1] - k.vis_off[f];
                if (V < 0 || V > N || nv < 0 || nv > N) bad = true;
                else {
                    const int* ext = k.ext + k.ext_off[f];
                    const int* vis = k.vis + k.vis_off[f];
                    for (int i = 0; i < (int)V; i++) if (ext[i] < 0 || ext[i] >= N || (i > 0 && ext[i] <= ext[i - 1])) bad = true;
                    for (int i = 0; i < (int)nv; i++) if (vis[i] < 0 || vis[i] >= N) bad = true;
                }
            }
            sm.bcast[1] = bad;
            if (bad) {
                fr.ctl[FC_STAGE] = 0; fr.ctl[FC_ITER] = 0; fr.ctl[FC_NN] = 0; fr.ctl[FC_NCHUNK] = 0; fr.ctl[FC_NPRI] = 0;
                fr.ctl[FC_STATUS] = ST_BAD_INPUT; fr.ctl[FC_STPRE] = 0;
            }
        }
        __syncthreads();
        const bool bad = sm.bcast[1] != 0;
        __syncthreads();
        if (bad) return A_FINISH_CALL;
    }
    int Nn, n_priors = 0, n_visible = 0;
    const double* priors = nullptr;
    const double* Hext = nullptr;
    int hstride = k.node_stride;
    if (k.mode == 0) {
        Nn = k.n_nodes ? k.n_nodes[f] : k.node_stride;
        const long long ys = (long long)f * k.node_stride;
        priors = k.priors ? k.priors + (long long)f * k.priors_stride * 4 : nullptr;
        n_priors = (k.priors && k.n_priors) ? k.n_priors[f] : 0;
        n_priors = n_priors < 0 ? 0 : (n_priors > k.priors_stride ? k.priors_stride : n_priors);    // never past the frame's rows
        n_visible = k.n_visible ? k.n_visible[f] : 0;
        Hext = k.H ? k.H + ys * k.node_stride : nullptr;
    } else {
        const int V = (int)(k.ext_off[f + 1] - k.ext_off[f]);
        if (stage == 0) {
            Nn = V;
            Hext = k.H ? k.H + (long long)f * k.node_stride * k.node_stride : nullptr;
            // guide nodes (trackdlo.cpp:913-921)
            const int* ext = k.ext + k.ext_off[f];
            const double* Yf = k.Y + (long long)f * k.node_stride * 3;
            double* guide = tq_yio(a, fr, 0);
            for (int i = tid; i < 3 * V; i += nt) {
                const int r = i / 3, d = i - 3 * r;
                guide[i] = (V != k.node_stride) ? Yf[ext[r] * 3 + d] : Yf[i];
            }
            __syncthreads();
        } else {
            Nn = k.node_stride;
            priors = tq_pri(a, fr);
            n_priors = __ldcg(fr.ctl + FC_NPRI);
            n_visible = V;
        }
    }
    const double* Yio = tq_yio(a, fr, stage);
    const int n_chunks = (int)((fr.m0 + a.chunk - 1) / a.chunk);
    if (tid == 0) {
        fr.ctl[FC_STAGE] = stage; fr.ctl[FC_ITER] = 0; fr.ctl[FC_NN] = Nn; fr.ctl[FC_NCHUNK] = n_chunks;
        fr.ctl[FC_USEVIS] = (n_visible != Nn) && (n_visible > 0) && (p.k_vis != 0);    // trackdlo.cpp:358
        fr.ctl[FC_NPRI] = n_priors; fr.ctl[FC_STATUS] = 0; fr.ctl[FC_NVIS] = n_visible;
        fr.ctl[FC_PHASE] = TK_PRUNE; fr.ctl[FC_PENDING] = n_chunks + 1;
        fr.scal[FS_SIGMA2] = k.sigma2[f];
    }
    if (Nn < 4) {                          // reference indexes rows 2 and Nn-3 (trackdlo.cpp:313-321)
        if (tid == 0) fr.ctl[FC_STATUS] = ST_TOO_FEW_NODES;
        __syncthreads();
        return A_FINISH_CALL;
    }
    // ---- Y0, current Y (node4), arc-length coordinates (trackdlo.cpp:203, 216-223)
    double* gY0 = scr + sc.Y0;
    double* gS = scr + sc.S;
    double* gN4 = scr + sc.NODE4;
    for (int i = tid; i < 3 * Nn; i += nt) { const double v = Yio[i]; sm.y0[i] = v; gY0[i] = v; }
    __syncthreads();
    for (int j = tid; j < Nn; j += nt) { gN4[4 * j] = sm.y0[3 * j]; gN4[4 * j + 1] = sm.y0[3 * j + 1]; gN4[4 * j + 2] = sm.y0[3 * j + 2]; gN4[4 * j + 3] = 0.0; }
    if (tid == 0) {
        double cur = 0.0;
        sm.s[0] = 0.0;
        for (int i = 0; i < Nn - 1; i++) {
            cur += sqrt(dist2(sm.y0[3 * i + 3], sm.y0[3 * i + 4], sm.y0[3 * i + 5], sm.y0[3 * i], sm.y0[3 * i + 1], sm.y0[3 * i + 2]));
            sm.s[i + 1] = cur;
        }
    }
    __syncthreads();
    for (int j = tid; j < Nn; j += nt) gS[j] = sm.s[j];
    // the PRUNE wave only needs node4: publish it now, do the rest of the set-up meanwhile.  A frame of a single chunk (a
    // live sequence with a few hundred to a few thousand points) never goes through the queue: this CTA runs the chunk
    // itself right after the set-up (one CTA owns the whole frame, no ticket / publish / pop latency per wave).
    if (n_chunks == 1) local_wd = tq_word(0, a.qmask, TK_PRUNE, f, 0);
    else tq_push(a, sm, TK_PRUNE, f, n_chunks);

    // ---- G, priors, LLE products (trackdlo.cpp:225-260).  The dense kernel matrix is only built when a dense solve
    // will use it (LLE regulariser, negative alpha, or the dense solver selected); the structured solve needs the
    // transitions of the Matern-3/2 state over the gaps h_t = s_{t+1} - s_t instead (mct_kalman_solve).
    double* gG = scr + sc.G;
    double* gHG = scr + sc.HG;
    double* gH = scr + sc.H;
    const double beta = p.beta;
    // 0 = structured (state-space filter) without LLE,
    // 1 = dense (also for a negative alpha, which the state-space form -- a square root of P1 + alpha J -- cannot take)
    // 3 = banded information-form solve with LLE (the default with LLE; needs distinct arc lengths: a zero gap has no
    // precision matrix -- such a chain falls back to the dense path)
    int smode = 1;
    if (a.solver != 1 && !(n_priors > 0 && p.alpha < 0.0)) {
        if (!p.include_lle) smode = 0;
        else if (!Hext) smode = 3;   // (also below 65 nodes: as fast as the register-resident elimination on an idle GPU, faster on a loaded one)
    }
    if (smode == 3) {
        int zero_gap = 0;
        for (int t = tid; t + 1 < Nn; t += nt) zero_gap |= !(sm.s[t + 1] - sm.s[t] > 1e-12 * beta);
        if (__syncthreads_or(zero_gap)) smode = 1;
    }
    const bool dense = smode == 1 || smode == 3;          // the dense kernel matrix is only built for smode 1 (below)
    if (tid == 0) fr.ctl[FC_DENSE] = smode;
    if (smode == 3) {
        // nothing here: K comes from the arc lengths after H is known (below)
    } else if (dense) {
        for (int idx = tid; idx < Nn * Nn; idx += nt) {
            const int i = idx / Nn, j = idx - i * Nn;
            const double dd = fabs(sm.s[i] - sm.s[j]);
            gG[idx] = 1 / (2 * beta * 2 * beta) * exp(-sqrt(2.0) * dd / beta) * (2 * dd + sqrt(2.0) * beta);
        }
    } else {
        // Phi(h) = e^{-ah} [[1 + ah, h], [-a^2 h, 1 - ah]],  Q(h) = P_inf - Phi P_inf Phi^T,  P_inf = sigma_f^2 diag(1, a^2)
        double* gPhi = scr + sc.PHI;
        const double ak = sqrt(2.0) / beta;
        for (int t = tid; t + 1 < Nn; t += nt) {
            const double h = fabs(sm.s[t + 1] - sm.s[t]), x = ak * h, e = exp(-x);
            const double p00 = e * (1.0 + x), p01 = e * h, p10 = -e * ak * x, p11 = e * (1.0 - x);
            gPhi[4 * t] = p00; gPhi[4 * t + 1] = p01; gPhi[4 * t + 2] = p10; gPhi[4 * t + 3] = p11;
        }
    }
    double* gJD = scr + sc.JD;
    double* gYE = scr + sc.YEXT;
    for (int i = tid; i < Nn; i += nt) gJD[i] = 0.0;
    for (int i = tid; i < 3 * Nn; i += nt) gYE[i] = sm.y0[i];
    __syncthreads();
    if (tid == 0) {
        for (int kk = 0; kk < n_priors; kk++) {
            const int idx = (int)priors[kk * 4];               // trackdlo.cpp:247
            if (idx < 0 || idx >= Nn) continue;
            gJD[idx] = 1.0;
            gYE[idx * 3] = priors[kk * 4 + 1]; gYE[idx * 3 + 1] = priors[kk * 4 + 2]; gYE[idx * 3 + 2] = priors[kk * 4 + 3];
        }
    }
    if (p.include_lle) {
        if (Hext) {
            for (int idx = tid; idx < Nn * Nn; idx += nt) { const int i = idx / Nn, j = idx - i * Nn; gH[idx] = Hext[(long long)i * hstride + j]; }
        } else {
            double* E = scr + sc.AB;                              // dense E = I - L
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
                for (int kk = klo; kk <= khi; kk++) s = __dadd_rn(s, __dmul_rn(E[(long long)kk * Nn + r], E[(long long)kk * Nn + c]));
                gH[idx] = s;
            }
        }
        __syncthreads();
        if (smode == 3) mct_banded_lle_setup(Nn, beta, p.lambda, p.gamma, sm.s, gH, scr + sc.KLW, scr + sc.KTR);
        else
        for (int idx = tid; idx < Nn * Nn; idx += nt) {            // HG = H G
            const int i = idx / Nn, j = idx - i * Nn;
            double s = 0.0;
            for (int kk = 0; kk < Nn; kk++) s = fma(gH[(long long)i * Nn + kk], gG[(long long)kk * Nn + j], s);
            gHG[idx] = s;
        }
        double* gHY = scr + sc.HY0;
        for (int idx = tid; idx < 3 * Nn; idx += nt) {             // H Y0
            const int i = idx / 3, d = idx - 3 * i;
            double s = 0.0;
            for (int kk = 0; kk < Nn; kk++) s = fma(gH[(long long)i * Nn + kk], sm.y0[3 * kk + d], s);
            gHY[idx] = s;
        }
    }
    return tq_arrive(sm, fr.ctl) ? A_AFTER_PRUNE : A_NONE;
}

// ------------------------------------------------------------------------------------------
// after_prune: gather the kept counts, sigma2 init (trackdlo.cpp:263-273)
// ------------------------------------------------------------------------------------------
static __device__ int tq_after_prune(const TqArgs& a, TqSm& sm, const TqFrame& fr) {
    const int tid = threadIdx.x;
    const int stage = __ldcg(fr.ctl + FC_STAGE), Nn = __ldcg(fr.ctl + FC_NN), n_chunks = __ldcg(fr.ctl + FC_NCHUNK);
    const CpdP& p = tq_params(a, stage);
    if (tid == 0) {
        long long Mp = 0;
        double sumd2 = 0.0;
        for (int c = 0; c < n_chunks; c++) { Mp += __ldcg(a.nkept + fr.gbase + c); sumd2 += __ldcg(a.gath + fr.gbase + c); }
        double sigma2 = __ldcg(fr.scal + FS_SIGMA2);
        if (Mp > 0 && sigma2 == 0) sigma2 = sumd2 / (3.0 * (double)Nn * (double)Mp);   // trackdlo.cpp:271-273
        fr.scal[FS_SIGMA2] = sigma2;
        fr.scal[FS_MP] = (double)Mp;
        sm.bcast[1] = Mp > 0;
    }
    __syncthreads();
    const bool nonempty = sm.bcast[1] != 0;
    __syncthreads();
    if (!nonempty) { if (tid == 0) fr.ctl[FC_STATUS] = ST_EMPTY; __syncthreads(); return A_FINISH_CALL; }
    if (p.max_iter <= 0) return A_FINISH_CALL;
    return A_BEGIN_ITER;
}

// ------------------------------------------------------------------------------------------
// begin_iter: per-iteration header (scaled arc lengths, outlier constant) and the next wave
// ------------------------------------------------------------------------------------------
static __device__ int tq_begin_iter(const TqArgs& a, TqSm& sm, const TqFrame& fr, unsigned long long& local_wd) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int stage = __ldcg(fr.ctl + FC_STAGE), Nn = __ldcg(fr.ctl + FC_NN), n_chunks = __ldcg(fr.ctl + FC_NCHUNK);
    const int use_vis = __ldcg(fr.ctl + FC_USEVIS);
    const CpdP& p = tq_params(a, stage);
    const double sigma2 = __ldcg(fr.scal + FS_SIGMA2), Mp = __ldcg(fr.scal + FS_MP);
    const double rscale = sqrt(0.5 / sigma2);
    double* gN4 = fr.scr + fr.sc.NODE4;
    const double* gS = fr.scr + fr.sc.S;
    for (int j = tid; j < Nn; j += nt) gN4[4 * j + 3] = __ldcg(gS + j) * rscale;
    if (tid == 0) {
        const double c_gauss = pow(2 * M_PI * sigma2, 1.5) * p.mu / (1 - p.mu);
        fr.scal[FS_RSCALE] = rscale;
        fr.scal[FS_CGAUSS] = c_gauss;
        fr.scal[FS_CNORM] = use_vis ? c_gauss / Mp : c_gauss * (double)Nn / Mp;       // trackdlo.cpp:378 / 300
        fr.ctl[FC_PHASE] = use_vis ? TK_DMIN : TK_ESTEP;
        fr.ctl[FC_PENDING] = n_chunks;
    }
    if (n_chunks == 1) { __syncthreads(); local_wd = tq_word(0, a.qmask, use_vis ? TK_DMIN : TK_ESTEP, fr.f, 0); }
    else tq_push(a, sm, use_vis ? TK_DMIN : TK_ESTEP, fr.f, n_chunks);
    return A_NONE;
}

// after_dmin: visibility weights (trackdlo.cpp:291-293, 358-375), then the E-step wave
static __device__ int tq_after_dmin(const TqArgs& a, TqSm& sm, const TqFrame& fr, unsigned long long& local_wd) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int stage = __ldcg(fr.ctl + FC_STAGE), Nn = __ldcg(fr.ctl + FC_NN), n_chunks = __ldcg(fr.ctl + FC_NCHUNK);
    const CpdP& p = tq_params(a, stage);
    const int N = a.k.scr_nodes;
    for (int j = tid; j < Nn; j += nt) {
        double m = 1e300;
        for (int c = 0; c < n_chunks; c++) m = fmin(m, __ldcg(a.dminp + (long long)(fr.gbase + c) * N + j));
        double dm = sqrt(m);
        if (dm <= p.tau) dm = 0.0;                              // trackdlo.cpp:291-293
        sm.vw[j] = exp(-p.k_vis * dm);                          // trackdlo.cpp:365
    }
    __syncthreads();
    if (tid == 0) { double t = 0.0; for (int j = 0; j < Nn; j++) t += sm.vw[j]; sm.red[40] = t; }
    __syncthreads();
    const double tot = sm.red[40];
    double* gVW = fr.scr + fr.sc.VW;
    for (int j = tid; j < Nn; j += nt) gVW[j] = sm.vw[j] / tot;      // trackdlo.cpp:372
    if (tid == 0) { fr.ctl[FC_PHASE] = TK_ESTEP; fr.ctl[FC_PENDING] = n_chunks; }
    if (n_chunks == 1) { __syncthreads(); local_wd = tq_word(0, a.qmask, TK_ESTEP, fr.f, 0); }
    else tq_push(a, sm, TK_ESTEP, fr.f, n_chunks);
    return A_NONE;
}

// ------------------------------------------------------------------------------------------
// m_step (trackdlo.cpp:392-438): run by the CTA that completed the frame's E-step wave.
// ------------------------------------------------------------------------------------------
static __device__ int tq_mstep(const TqArgs& a, TqSm& sm, const TqFrame& fr, long long& tprev) {
    unsigned long long* prof = a.k.prof;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5, nt = blockDim.x;
    const int stage = __ldcg(fr.ctl + FC_STAGE), Nn = __ldcg(fr.ctl + FC_NN), n_chunks = __ldcg(fr.ctl + FC_NCHUNK);
    const int it = __ldcg(fr.ctl + FC_ITER);
    int status = __ldcg(fr.ctl + FC_STATUS);
    const bool have_priors = __ldcg(fr.ctl + FC_NPRI) > 0;
    const CpdP& p = tq_params(a, stage);
    const double sigma2 = __ldcg(fr.scal + FS_SIGMA2);
    const TqScr& sc = fr.sc;
    double* scr = fr.scr;
    const double* gG = scr + sc.G;
    const double* gHG = scr + sc.HG;
    const int ld = Nn + 3;
    const int smode = __ldcg(fr.ctl + FC_DENSE);
    const bool dense = smode == 1;                          // 0 / 2 / 3: structured solves (no dense matrix)
    // [A|B] in shared memory only for the register-resident solvers (Nn <= 64); the blocked Cholesky keeps it in global
    // scratch (its shared workspace reuses the whole tail of the region)
    const bool ab_in_smem = Nn <= 64 && (long long)Nn * ld <= (long long)a.L.ab_doubles;
    double* AB = ab_in_smem ? sm.ab : scr + sc.AB;

    // ---- frame vectors -> shared; partial sums in chunk order.  The M-step is a latency chain on the frame's
    // critical path: all loads of this block are requested before any is consumed (two L2 round trips in total).
    {
        constexpr int CB = 20;                                    // chunks gathered per batch
        const int np1 = 4 * Nn + 1;                               // [Nn][4] sums + sum Pt1 |x|^2
        const double* src = a.part + (long long)fr.gbase * a.part_stride + tid;
        double t[CB];
        const bool gth = tid < np1;
#pragma unroll
        for (int u = 0; u < CB; u++) t[u] = (gth && u < n_chunks) ? __ldcg(src + (long long)u * a.part_stride) : 0.0;
        double vy[3], ve[3], vh[3];                               // 3 Nn <= 3 nt for Nn <= nt (else looped below)
#pragma unroll
        for (int u = 0; u < 3; u++) {
            const int i = tid + u * nt;
            const bool ok = i < 3 * Nn;
            vy[u] = ok ? __ldcg(scr + sc.Y0 + i) : 0.0;
            ve[u] = (ok && have_priors) ? __ldcg(scr + sc.YEXT + i) : 0.0;
            vh[u] = (ok && p.include_lle) ? __ldcg(scr + sc.HY0 + i) : 0.0;
        }
        const bool nd_ok = tid < Nn;
        const double vj = (nd_ok && have_priors) ? __ldcg(scr + sc.JD + tid) : 0.0;
        double4 q4 = make_double4(0.0, 0.0, 0.0, 0.0);
        if (nd_ok) q4 = ldcg4(reinterpret_cast<const double4*>(scr + sc.NODE4) + tid);
        // consume
#pragma unroll
        for (int u = 0; u < 3; u++) {
            const int i = tid + u * nt;
            if (i < 3 * Nn) { sm.y0[i] = vy[u]; sm.yext[i] = ve[u]; sm.hy0[i] = vh[u]; }
        }
        if (nd_ok) { sm.jd[tid] = vj; sm.node4[tid] = q4; }
        for (int i = tid + 3 * nt; i < 3 * Nn; i += nt) {         // Nn > nt (not reachable with Nn <= 256, nt >= 224 ... kept for safety)
            sm.y0[i] = __ldcg(scr + sc.Y0 + i);
            sm.yext[i] = have_priors ? __ldcg(scr + sc.YEXT + i) : 0.0;
            sm.hy0[i] = p.include_lle ? __ldcg(scr + sc.HY0 + i) : 0.0;
        }
        for (int i = tid + nt; i < Nn; i += nt) {
            sm.jd[i] = have_priors ? __ldcg(scr + sc.JD + i) : 0.0;
            sm.node4[i] = ldcg4(reinterpret_cast<const double4*>(scr + sc.NODE4) + i);
        }
        double v = 0.0;
#pragma unroll
        for (int u = 0; u < CB; u++) v += t[u];                   // chunk order (entries beyond n_chunks are +0.0)
        for (int c0 = CB; c0 < n_chunks; c0 += CB) {
#pragma unroll
            for (int u = 0; u < CB; u++) t[u] = (gth && c0 + u < n_chunks) ? __ldcg(src + (long long)(c0 + u) * a.part_stride) : 0.0;
#pragma unroll
            for (int u = 0; u < CB; u++) v += t[u];
        }
        if (gth) {
            if (tid == 4 * Nn) sm.red[42] = v;
            else { const int m = tid >> 2, kk = tid & 3; if (kk == 0) sm.p1[m] = v; else sm.px[3 * m + kk - 1] = v; }
        }
        for (int i = tid + nt; i < np1; i += nt) {                // 4 Nn + 1 > nt (Nn > 55 with 224 threads)
            const double* s2 = a.part + (long long)fr.gbase * a.part_stride + i;
            double v2 = 0.0;
            for (int c = 0; c < n_chunks; c++) v2 += __ldcg(s2 + (long long)c * a.part_stride);
            if (i == 4 * Nn) sm.red[42] = v2;
            else { const int m = i >> 2, kk = i & 3; if (kk == 0) sm.p1[m] = v2; else sm.px[3 * m + kk - 1] = v2; }
        }
    }
    // G -> shared (behind [A|B]) when it fits: used by the assembly and by T = Y0 + G W
    const bool g_in_smem = dense && ab_in_smem && (long long)Nn * ld + (long long)Nn * Nn <= (long long)a.L.ab_doubles;
    double* sG = sm.ab + Nn * ld;
    if (g_in_smem) {
        double gv[20];                                            // Nn^2 <= 4096 <= 20 * 224
#pragma unroll
        for (int u = 0; u < 20; u++) { const int idx = tid + u * nt; gv[u] = idx < Nn * Nn ? __ldcg(gG + idx) : 0.0; }
#pragma unroll
        for (int u = 0; u < 20; u++) { const int idx = tid + u * nt; if (idx < Nn * Nn) sG[idx] = gv[u]; }
    }
    __syncthreads();
    const double sxx = sm.red[42];
    int sing;
    if (!dense) {
        // ---- structured path (no LLE): D = P1 + alpha J, B = PX - P1 Y0 + alpha (Yext - Y0) (trackdlo.cpp:407-412), then the
        // O(Nn) state-space solve for W and T = Y0 + G W.  The workspace overlays yext / jd / hy0 / [A|B]: everything read
        // from there goes through registers first.
        double bv[3], dv = 0.0;
#pragma unroll
        for (int u = 0; u < 3; u++) {
            const int idx = tid + u * nt;
            bv[u] = 0.0;
            if (idx < 3 * Nn) {
                const int i = idx / 3;
                bv[u] = sm.px[idx] - sm.p1[i] * sm.y0[idx] + (have_priors ? p.alpha * (sm.yext[idx] - sm.y0[idx]) : 0.0);
                if (smode == 3) bv[u] -= sigma2 * p.gamma * sm.hy0[idx];          // - sigma2 gamma H Y0 (trackdlo.cpp:400)
            }
        }
        if (tid < Nn) dv = sm.p1[tid] + (have_priors ? p.alpha * sm.jd[tid] : 0.0);
        double4 ph4 = make_double4(0.0, 0.0, 0.0, 0.0);
        if (smode == 0 && tid + 1 < Nn) ph4 = ldcg4(reinterpret_cast<const double4*>(scr + sc.PHI) + tid);
        __syncthreads();
        double* kdd = sm.yext;              // [Nn]
        double* kbt = kdd + Nn;             // [3][Nn]
        if (smode == 3) {
            TQ_TICK(6)
            double* bw = sm.gjbuf;
            bw += (reinterpret_cast<uintptr_t>(bw) >> 3) & 1;          // 16-byte aligned
            sing = mct_banded_lle_solve(Nn, sigma2, dv, bv, scr + sc.KLW, scr + sc.KTR, sm.y0, bw, sm.wsol, sm.tnew);
            if (sing) status |= ST_SINGULAR;
            TQ_TICK(7)
        } else {
        double* kphi = kbt + 3 * Nn;        // [Nn][4]
        double* krf = kphi + 4 * Nn;        // [Nn]
        double* kkk = krf + Nn;             // [2 Nn]
        double* kpp = kkk + 2 * Nn;         // [2 Nn]
        double* kam = kpp + 2 * Nn;         // [6 Nn]
#pragma unroll
        for (int u = 0; u < 3; u++) {
            const int idx = tid + u * nt;
            if (idx < 3 * Nn) { const int i = idx / 3, d = idx - 3 * i; kbt[d * Nn + i] = bv[u]; }
        }
        if (tid < Nn) { kdd[tid] = dv; *reinterpret_cast<double4*>(kphi + 4 * tid) = ph4; }
        for (int i = tid + nt; i < Nn; i += nt) {            // (Nn > nt: not reachable with Nn <= 256)
            kdd[i] = sm.p1[i] + (have_priors ? p.alpha * sm.jd[i] : 0.0);
            *reinterpret_cast<double4*>(kphi + 4 * i) = ldcg4(reinterpret_cast<const double4*>(scr + sc.PHI) + i);
        }
        __syncthreads();
        TQ_TICK(6)
        sing = mct_kalman_solve(Nn, p.lambda * sigma2, p.beta, kdd, kbt, kphi, sm.y0, sm.wsol, sm.tnew, krf, kkk, kpp, kam);
        if (sing) status |= ST_SINGULAR;
        TQ_TICK(7)
        }
    } else {

    // ---- assemble [A | B] (trackdlo.cpp:392-413).  Without LLE, A = D'G + cI (D' = diag(P1 + alpha J), c = lambda sigma2)
    // is similar to the SPD matrix D'^1/2 G D'^1/2 + cI: solve that for Z = D'^-1/2 W (rows with D' = 0 have B = 0)
    const double ls = p.lambda * sigma2, sg = sigma2 * p.gamma;
    const bool small = ab_in_smem && Nn <= 64;
    const int chol_nb = (!small && !p.include_lle && !(have_priors && p.alpha < 0.0)) ? ((long long)Nn * 20 + Nn + 64 <= (long long)a.L.chol_doubles ? 16 : ((long long)Nn * 12 + Nn + 64 <= (long long)a.L.chol_doubles ? 8 : 0)) : 0;
    // (a negative alpha would put a negative number under the square root: the reference's generic solve accepts it, so
    // does the pivoted path here)
    const bool spd = !p.include_lle && !(have_priors && p.alpha < 0.0) && (small || chol_nb > 0);
    if (spd) {
        for (int i = tid; i < Nn; i += nt) sm.tnew[i] = sqrt(sm.p1[i] + (have_priors ? p.alpha * sm.jd[i] : 0.0));
        __syncthreads();
    }
    if (g_in_smem && !spd) {
        // Nn <= 64, pivoted path (LLE): all H G entries of this thread are requested from L2 before the first is used
        double hv[20];                                            // Nn^2 <= 4096 <= 20 * 224
#pragma unroll
        for (int u = 0; u < 20; u++) { const int idx = tid + u * nt; hv[u] = (p.include_lle && idx < Nn * Nn) ? __ldcg(gHG + idx) : 0.0; }
#pragma unroll
        for (int u = 0; u < 20; u++) {
            const int idx = tid + u * nt;
            if (idx < Nn * Nn) {
                const int i = idx / Nn, j = idx - i * Nn;
                const double g = sG[idx];
                double v = sm.p1[i] * g + (i == j ? ls : 0.0);
                if (p.include_lle) v += sg * hv[u];
                if (have_priors) v += p.alpha * sm.jd[i] * g;
                AB[(long long)i * ld + j] = v;
            }
        }
    } else
    for (int idx = tid; idx < Nn * Nn; idx += nt) {
        const int i = idx / Nn, j = idx - i * Nn;
        const double g = g_in_smem ? sG[idx] : __ldcg(gG + idx);
        double v;
        if (spd) v = sm.tnew[i] * g * sm.tnew[j] + (i == j ? ls : 0.0);
        else {
            v = sm.p1[i] * g + (i == j ? ls : 0.0);
            if (p.include_lle) v += sg * __ldcg(gHG + idx);
            if (have_priors) v += p.alpha * sm.jd[i] * g;
        }
        AB[(long long)i * ld + j] = v;
    }
    for (int idx = tid; idx < 3 * Nn; idx += nt) {
        const int i = idx / 3;
        double v = sm.px[idx] - sm.p1[i] * sm.y0[idx];
        if (p.include_lle) v -= sg * sm.hy0[idx];
        if (have_priors) v += p.alpha * (sm.yext[idx] - sm.y0[idx]);
        if (spd) { const double sd = sm.tnew[i]; v = sd > 0.0 ? v / sd : 0.0; }
        AB[(long long)i * ld + Nn + (idx - 3 * i)] = v;
    }
    __syncthreads();
    TQ_TICK(6)
    if (small) {
        double sdreg[3];
        if (spd) for (int t = 0, i = tid; t < 3 && i < 3 * Nn; t++, i += nt) sdreg[t] = sm.tnew[i / 3];
        // register-resident elimination when the warps cover the columns, else the shared-memory one
        if (nw * 8 >= ld) sing = spd ? gj_solve_regs<8, false>(sm.ab, Nn, ld, sm.gjbuf, sm.wsol) : gj_solve_regs<8, true>(sm.ab, Nn, ld, sm.gjbuf, sm.wsol);
        else if (nw * 10 >= ld) sing = spd ? gj_solve_regs<10, false>(sm.ab, Nn, ld, sm.gjbuf, sm.wsol) : gj_solve_regs<10, true>(sm.ab, Nn, ld, sm.gjbuf, sm.wsol);
        else sing = gj_solve_small(sm.ab, Nn, ld, sm.gjbuf, sm.prow, sm.wsol, !spd, nullptr);
        if (spd) {
            for (int t = 0, i = tid; t < 3 && i < 3 * Nn; t++, i += nt) sm.wsol[i] *= sdreg[t];
            __syncthreads();
        }
    } else if (chol_nb > 0) {
        // Nn > 64, SPD form: blocked Cholesky with FP64 tensor-core trailing updates; rhs = B columns of [A|B]
        double sdreg[3];
        for (int t = 0, i = tid; t < 3 && i < 3 * Nn; t++, i += nt) { sdreg[t] = sm.tnew[i / 3]; sm.wsol[i] = AB[(long long)(i / 3) * ld + Nn + (i % 3)]; }
        __syncthreads();
        sing = chol_nb == 16 ? chol_solve_blocked<16>(AB, Nn, ld, sm.yext, sm.wsol) : chol_solve_blocked<8>(AB, Nn, ld, sm.yext, sm.wsol);   // workspace: yext .. end of the region
        for (int t = 0, i = tid; t < 3 && i < 3 * Nn; t++, i += nt) sm.wsol[i] *= sdreg[t];
        __syncthreads();
    } else sing = gj_solve(AB, Nn, ld, sm.prow, sm.used, sm.red + 43, sm.wsol);
    if (sing) status |= ST_SINGULAR;
    TQ_TICK(7)

    // ---- T = Y0 + G W (trackdlo.cpp:417)
    if (g_in_smem) {
        for (int i = tid; i < 3 * Nn; i += nt) {
            const int r = i / 3, d = i - 3 * r;
            const double* __restrict__ grow = sG + r * Nn;
            double acc = 0.0;
            for (int kk = 0; kk < Nn; kk++) acc = fma(grow[kk], sm.wsol[3 * kk + d], acc);
            sm.tnew[i] = sm.y0[i] + acc;
        }
    } else {
        for (int i = warp; i < Nn; i += nw) {
            double ax = 0.0, ay = 0.0, az = 0.0;
            for (int kk = lane; kk < Nn; kk += 32) {
                const double g = __ldcg(gG + (long long)i * Nn + kk);
                ax = fma(g, sm.wsol[3 * kk], ax); ay = fma(g, sm.wsol[3 * kk + 1], ay); az = fma(g, sm.wsol[3 * kk + 2], az);
            }
            ax = warp_sum(ax); ay = warp_sum(ay); az = warp_sum(az);
            if (lane == 0) { sm.tnew[3 * i] = sm.y0[3 * i] + ax; sm.tnew[3 * i + 1] = sm.y0[3 * i + 1] + ay; sm.tnew[3 * i + 2] = sm.y0[3 * i + 2] + az; }
        }
    }
    }       // dense path
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
            fr.scal[FS_SIGMA2] = s2new;
            fr.ctl[FC_STATUS] = status;
            fr.ctl[FC_ITER] = it + 1;
            sm.bcast[1] = fin;
        }
    }
    double* gN4 = scr + sc.NODE4;
    for (int j = tid; j < Nn; j += nt) { gN4[4 * j] = sm.tnew[3 * j]; gN4[4 * j + 1] = sm.tnew[3 * j + 1]; gN4[4 * j + 2] = sm.tnew[3 * j + 2]; }
    for (int i = tid; i < 3 * Nn; i += nt) scr[sc.WSOL + i] = sm.wsol[i];
    __syncthreads();
    const int fin = sm.bcast[1];
    __syncthreads();
    TQ_TICK(8)
    return fin ? A_FINISH_CALL : A_BEGIN_ITER;
}

// ------------------------------------------------------------------------------------------
// finish_call: results of one cpd_lle call; in tracking mode the glue between the pre-processing and the
// main registration (trackdlo.cpp:929-998).
// ------------------------------------------------------------------------------------------
static __device__ int tq_finish_call(const TqArgs& a, TqSm& sm, const TqFrame& fr, int& next_stage) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const KArgs& k = a.k;
    const int f = fr.f;
    const int stage = __ldcg(fr.ctl + FC_STAGE), Nn = __ldcg(fr.ctl + FC_NN);
    const int status = __ldcg(fr.ctl + FC_STATUS), iters = __ldcg(fr.ctl + FC_ITER);
    const CpdP& p = tq_params(a, stage);
    const bool ran = !(status & (ST_TOO_FEW_NODES | ST_EMPTY | ST_BAD_INPUT));
    double* Yio = tq_yio(a, fr, stage);
    const double* gN4 = fr.scr + fr.sc.NODE4;
    if (ran) for (int i = tid; i < 3 * Nn; i += nt) Yio[i] = __ldcg(gN4 + 4 * (i / 3) + (i % 3));
    if (k.mode == 0) {
        if (ran && k.W && p.max_iter > 0) for (int i = tid; i < 3 * Nn; i += nt) k.W[(long long)f * k.node_stride * 3 + i] = __ldcg(fr.scr + fr.sc.WSOL + i);
        if (tid == 0) {
            if (ran) k.sigma2[f] = __ldcg(fr.scal + FS_SIGMA2);
            if (k.iters) k.iters[f] = ran ? iters : 0;
            if (k.status) k.status[f] = status;
        }
        return A_FRAME_DONE;
    }
    const int N = k.node_stride;
    if (stage == 0) {
        if (tid == 0) { if (k.iters) k.iters[2 * f] = ran ? iters : 0; fr.ctl[FC_ITPRE] = ran ? iters : 0; }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            int st = 0;
            if (status & ST_NOT_CONVERGED) st |= ST_PRE_NOT_CONVERGED;
            st |= status & ~ST_NOT_CONVERGED;
            const int* vis = k.vis + k.vis_off[f];
            const int nvis = (int)(k.vis_off[f + 1] - k.vis_off[f]);
            const int* ext = k.ext + k.ext_off[f];
            const int V = Nn;
            const double* guide = Yio;
            const double* Yf = k.Y + (long long)f * N * 3;
            const double* geo = k.rest + (long long)f * N;
            double* pri = tq_pri(a, fr);
            int err = 0, state = 0, np = 0;
            double* trv = fr.scr + fr.sc.TRV;
            if (!ran) {
                state = -1;
            } else if (V == N) {
                state = 0;
                double* v1 = trv;
                double* v2 = trv + (N + 2) * 4;
                const int n1 = traverse_euclidean(geo, N, guide, V, ext, V, 0, -1, v1, &err);
                const int n2 = traverse_euclidean(geo, N, guide, V, ext, V, 1, -1, v2, &err);
                // v2 is emitted tail -> head; the reference reverses it (trackdlo.cpp:942): v2r[j] = v2[n2-1-j]
                for (int i = 0; i < N; i++) {
                    const int j2 = i - (N - n2);
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
            } else if (ext[0] == 0 && ext[V - 1] == N - 1) {
                state = 1;
                np = traverse_euclidean(geo, N, guide, V, ext, V, 0, -1, pri, &err);
                np += traverse_euclidean(geo, N, guide, V, ext, V, 1, -1, pri + np * 4, &err);
            } else if (ext[0] == 0) {
                state = 2;
                np = traverse_euclidean(geo, N, guide, V, ext, V, 0, -1, pri, &err);
            } else if (ext[V - 1] == N - 1) {
                state = 3;
                np = traverse_euclidean(geo, N, guide, V, ext, V, 1, -1, pri, &err);
            } else {
                state = 4;
                int align = -1;
                double moved = 999999;
                for (int i = 0; i < nvis; i++) {
                    if (i >= V) { err |= 8; break; }
                    const double dd = vdist(ld3(Yf, vis[i]), ld3(guide, i));
                    if (dd < moved) { moved = dd; align = i; }
                }
                np = traverse_euclidean(geo, N, guide, V, ext, V, 2, align, pri, &err);
            }
            if (err) st |= ST_TRAVERSE_UB;
            if (k.state_out) k.state_out[f] = state;
            if (k.n_priors_out) k.n_priors_out[f] = np;
            fr.ctl[FC_NPRI] = np;
            fr.ctl[FC_STPRE] = st;
            sm.bcast[1] = ran;
            __threadfence();
        }
        __syncthreads();
        const bool go = sm.bcast[1] != 0;
        __syncthreads();
        if (go) { next_stage = 1; return A_START_CALL; }
        if (tid == 0) { if (k.status) k.status[f] = __ldcg(fr.ctl + FC_STPRE); if (k.iters) k.iters[2 * f + 1] = 0; }
        if (k.packed_out) {                 // frame not run: nodes unchanged
            double* rec = k.packed_out + (long long)f * (3 * N + 4);
            const double* Yf = k.Y + (long long)f * N * 3;
            for (int i = tid; i < 3 * N; i += nt) rec[i] = Yf[i];
            if (tid == 0) { rec[3 * N] = k.sigma2[f]; rec[3 * N + 1] = 0.0; rec[3 * N + 2] = 0.0; rec[3 * N + 3] = (double)__ldcg(fr.ctl + FC_STPRE); }
        }
        return A_FRAME_DONE;
    }
    // stage 1: main registration done
    if (ran && k.W && p.max_iter > 0) for (int i = tid; i < 3 * Nn; i += nt) k.W[(long long)f * N * 3 + i] = __ldcg(fr.scr + fr.sc.WSOL + i);
    if (tid == 0) {
        if (ran) k.sigma2[f] = __ldcg(fr.scal + FS_SIGMA2);
        if (k.iters) k.iters[2 * f + 1] = ran ? iters : 0;
        if (k.status) k.status[f] = status | __ldcg(fr.ctl + FC_STPRE);
    }
    if (k.packed_out) {                     // one contiguous record per frame: what a multi-GPU caller all-gathers
        double* rec = k.packed_out + (long long)f * (3 * N + 4);
        const double* Yf = k.Y + (long long)f * N * 3;
        for (int i = tid; i < 3 * N; i += nt) rec[i] = ran ? __ldcg(gN4 + 4 * (i / 3) + (i % 3)) : Yf[i];
        if (tid == 0) {
            rec[3 * N] = ran ? __ldcg(fr.scal + FS_SIGMA2) : k.sigma2[f];
            rec[3 * N + 1] = (double)__ldcg(fr.ctl + FC_ITPRE); rec[3 * N + 2] = ran ? (double)iters : 0.0;
            rec[3 * N + 3] = (double)(status | __ldcg(fr.ctl + FC_STPRE));
        }
    }
    return A_FRAME_DONE;
}

// ------------------------------------------------------------------------------------------
// The persistent kernel
// ------------------------------------------------------------------------------------------
// THREADS x MINB: 256 x 2 (128 registers, 16 warps/SM).  The shared-memory
// layout is a compile-time constant (sized for 32*NPASS nodes) so that no address arithmetic survives in the loops.
template <int NPASS, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) tdlo_tq_kernel(const TqArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    constexpr int nt = THREADS;
    constexpr TqSmemL L = tq_smem_layout(32 * NPASS, THREADS / 32);
    TqSm sm;
    sm.tab = reinterpret_cast<double*>(smem_raw + L.tab);
    sm.node4 = reinterpret_cast<double4*>(smem_raw + L.node4);
    sm.nsoa = reinterpret_cast<double*>(smem_raw + L.nsoa);
    sm.vw = reinterpret_cast<double*>(smem_raw + L.vw);
    sm.bcast = reinterpret_cast<int*>(smem_raw + L.bcast);
    sm.red = reinterpret_cast<double*>(smem_raw + L.red);
    sm.wbuf = reinterpret_cast<double4*>(smem_raw + L.wbuf);
    sm.ptile = reinterpret_cast<double*>(smem_raw + L.ptile);
    sm.wacc = reinterpret_cast<double*>(smem_raw + L.wacc);
    sm.y0 = reinterpret_cast<double*>(smem_raw + L.y0);
    sm.s = reinterpret_cast<double*>(smem_raw + L.s);
    sm.yext = reinterpret_cast<double*>(smem_raw + L.yext);
    sm.jd = reinterpret_cast<double*>(smem_raw + L.jd);
    sm.hy0 = reinterpret_cast<double*>(smem_raw + L.hy0);
    sm.p1 = reinterpret_cast<double*>(smem_raw + L.p1);
    sm.px = reinterpret_cast<double*>(smem_raw + L.px);
    sm.wsol = reinterpret_cast<double*>(smem_raw + L.wsol);
    sm.tnew = reinterpret_cast<double*>(smem_raw + L.tnew);
    sm.gjbuf = reinterpret_cast<double*>(smem_raw + L.gjbuf);
    sm.prow = reinterpret_cast<int*>(smem_raw + L.prow);
    sm.used = reinterpret_cast<int*>(smem_raw + L.used);
    sm.ab = reinterpret_cast<double*>(smem_raw + L.ab);
    if (NPASS <= 2) { for (int i = tid; i < EXP_TAB; i += nt) sm.tab[i] = a.k.exp_tab[i]; }            // 2^(i/2048)
    else { for (int i = tid; i < 64; i += nt) sm.tab[i] = a.k.exp_tab[i * (EXP_TAB / 64)]; }           // 2^(i/64)
    __syncthreads();

    int* qi = reinterpret_cast<int*>(a.qctl + 2);          // [0] next frame, [1] start tickets, [2] frames done, [3] abort flag
    unsigned long long* prof = a.k.prof;
    long long tprev = prof ? clock64() : 0;
    int action = A_NONE, af = 0, astage = 0;
    unsigned long long local_wd = 0;       // a task this CTA hands to itself (single-chunk frames)
    // The first `inflight` CTAs to get here each CLAIM a frame from the shared counter (not "frame = blockIdx.x"): a CTA
    // that becomes resident late (MPS, a debugger, a shared GPU) can then never restart a frame somebody else already ran.
    if (tid == 0) {
        int nf = -1;
        if (atomicAdd(qi + 1, 1) < a.inflight) { nf = atomicAdd(qi, 1); if (nf >= a.k.n_frames) nf = -1; }
        sm.bcast[1] = nf;
    }
    __syncthreads();
    { const int nf = sm.bcast[1]; if (nf >= 0) { action = A_START_CALL; af = nf; astage = 0; } }
    __syncthreads();

    for (;;) {
        // ---- continuations owned by this CTA
        while (action != A_NONE) {
            const TqFrame fr = tq_frame(a, af);
            switch (action) {
                case A_START_CALL: action = tq_start_call(a, sm, fr, astage, local_wd); TQ_TICK(4) break;
                case A_AFTER_PRUNE: action = tq_after_prune(a, sm, fr); TQ_TICK(5) break;
                case A_BEGIN_ITER: action = tq_begin_iter(a, sm, fr, local_wd); TQ_TICK(5) break;
                case A_AFTER_DMIN: action = tq_after_dmin(a, sm, fr, local_wd); TQ_TICK(5) break;
                case A_MSTEP: action = tq_mstep(a, sm, fr, tprev); break;
                case A_FINISH_CALL: action = tq_finish_call(a, sm, fr, astage); TQ_TICK(9) break;
                case A_FRAME_DONE: {
                    __threadfence();
                    __syncthreads();
                    if (tid == 0) {
                        const int nf = atomicAdd(qi, 1);
                        const int done = atomicAdd(qi + 2, 1) + 1;
                        sm.bcast[1] = nf < a.k.n_frames ? nf : -1;
                        sm.bcast[2] = done == a.k.n_frames;
                    }
                    __syncthreads();
                    const int nf = sm.bcast[1];
                    const bool all = sm.bcast[2] != 0;
                    __syncthreads();
                    if (all) tq_push(a, sm, TK_EXIT, 0, gridDim.x);
                    if (nf >= 0) { action = A_START_CALL; af = nf; astage = 0; }
                    else action = A_NONE;
                    break;
                }
                default: action = A_NONE;
            }
        }
        // ---- next task
        const unsigned long long wd = local_wd ? local_wd : tq_pop(a, sm);
        local_wd = 0;
        TQ_TICK(0)
        const int type = (int)((wd >> 37) & 7), f = (int)((wd >> 20) & 0x1ffff), c = (int)(wd & 0xfffff);
        if (type == TK_EXIT) break;
        const TqFrame fr = tq_frame(a, f);
        const int g = fr.gbase + c;
        const long long r0 = (long long)c * a.chunk;
        const long long r1 = r0 + a.chunk < fr.m0 ? r0 + a.chunk : fr.m0;
        // everything the task needs from L2 is requested in one go (the loads are independent: one round trip)
        const int Nn = __ldcg(fr.ctl + FC_NN), stage = __ldcg(fr.ctl + FC_STAGE), use_vis = __ldcg(fr.ctl + FC_USEVIS);
        const double sigma2 = __ldcg(fr.scal + FS_SIGMA2), c_norm = __ldcg(fr.scal + FS_CNORM), rscale = __ldcg(fr.scal + FS_RSCALE);
        const int n_kept = type == TK_PRUNE ? 0 : __ldcg(a.nkept + g);
        for (int j = tid; j < a.k.scr_nodes; j += nt) {            // rows beyond Nn are scratch: loaded, never used
            const double4 q = ldcg4(reinterpret_cast<const double4*>(fr.scr + fr.sc.NODE4) + j);
            const double v = __ldcg(fr.scr + fr.sc.VW + j);
            if (j < Nn) {
                sm.node4[j] = q; sm.vw[j] = v;
                sm.nsoa[j] = q.w;
                float* nf = reinterpret_cast<float*>(sm.nsoa + 32 * NPASS);
                nf[j] = (float)q.x; nf[32 * NPASS + j] = (float)q.y; nf[64 * NPASS + j] = (float)q.z;
            }
        }
        __syncthreads();
        if (type == TK_PRUNE) {
            Smem os;
            os.node4 = sm.node4; os.ptile = sm.ptile; os.red = sm.red;
            double sum_local;
            const int kept = prune_sort_slice(os, fr.Xraw, r0, r1, fr.Xc, a.k.bkt + (fr.Xraw - a.k.X) / 3, Nn,
                                              tq_params(a, stage).prune_radius, &sum_local);
            if (tid == 0) { __stcg(a.nkept + g, kept); __stcg(a.gath + g, sum_local); }
            __syncthreads();                                      // the sorted points of this chunk are in place
            tq_tile_spheres(fr.Xc + r0 * 3, kept, a.tsph + (long long)g * (a.chunk >> 5));
        } else if (type == TK_DMIN) {
            tq_dmin_chunk<NPASS>(sm, fr.Xc + r0 * 3, n_kept, Nn, a.dminp + (long long)g * a.k.scr_nodes);
        } else {
            double* part = a.part + (long long)g * a.part_stride;
            if (prof && tid == 0) atomicAdd(prof + 10, 1ull);
            if (use_vis) tq_estep_chunk<NPASS, true, THREADS / 32>(sm, fr.Xc + r0 * 3, a.tsph + (long long)g * (a.chunk >> 5), n_kept, Nn, sigma2, c_norm, rscale, a.zcut, a.zrel, tq_params(a, stage).k_vis, part, prof);
            else tq_estep_chunk<NPASS, false, THREADS / 32>(sm, fr.Xc + r0 * 3, a.tsph + (long long)g * (a.chunk >> 5), n_kept, Nn, sigma2, c_norm, rscale, a.zcut, a.zrel, tq_params(a, stage).k_vis, part, prof);
        }
        TQ_TICK(type)
        if (tq_arrive(sm, fr.ctl)) {
            af = f;
            action = type == TK_PRUNE ? A_AFTER_PRUNE : (type == TK_DMIN ? A_AFTER_DMIN : A_MSTEP);
        }
    }
}

}  // namespace tdlo
