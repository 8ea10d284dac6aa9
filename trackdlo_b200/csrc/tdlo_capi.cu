// Host side of the C ABI declared in include/trackdlo_b200.h.
// Replaces the host-visible surface of trackdlo::cpd_lle / trackdlo::tracking_step
// (trackdlo/include/trackdlo.h:81-102) for batches of independent frames.
#include "../../include/trackdlo_b200.h"
#include "tdlo_taskq.cuh"
#include "tdlo_visibility.cuh"
#include "tdlo_frontend.cuh"

#include <vector>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <dlfcn.h>

using namespace tdlo;

struct tdlo_ctx {
    int device = 0;
    int sm_count = 0;
    int max_frames = 0, max_nodes = 0;
    long long max_points = 0;
    cudaStream_t stream = nullptr;
    // pipelined upload of the point clouds (host-buffer entry points, task-queue engine): copy stream + progress flag
    cudaStream_t copy_stream = nullptr; cudaEvent_t ev_small = nullptr, ev_copy = nullptr; int* d_ready = nullptr; int* h_ready_vals = nullptr;
    const int* cur_ready = nullptr; int cur_ready_frames = 0;
    // device staging for the host-pointer entry points
    double *d_X = nullptr, *d_Y = nullptr, *d_sigma2 = nullptr, *d_priors = nullptr, *d_H = nullptr, *d_W = nullptr;
    double *d_rest = nullptr, *d_guide = nullptr, *d_priors_out = nullptr, *d_packed = nullptr;
    size_t priors_cap = 0;          // doubles in d_priors
    long long *d_xoff = nullptr, *d_visoff = nullptr, *d_extoff = nullptr;
    int *d_nnodes = nullptr, *d_npriors = nullptr, *d_nvis = nullptr, *d_iters = nullptr, *d_status = nullptr;
    int *d_vis = nullptr, *d_ext = nullptr, *d_npri_out = nullptr, *d_state = nullptr;
    // workspace
    double* d_Xc = nullptr;
    unsigned short* d_bkt = nullptr;
    double* d_exp_tab = nullptr;    // 2^(j/2048), j = 0..2047
    unsigned long long* d_prof = nullptr;   // phase cycle counters (enabled by tdlo_profile_phases)
    unsigned long long* d_prof_buf = nullptr;
    // visibility front-end workspace
    unsigned long long* d_vbits = nullptr; int *d_vtmp = nullptr, *d_vcnt = nullptr; long long* d_vslice = nullptr;
    int* d_vfree = nullptr; double* d_vproj = nullptr;      // self-occlusion test: flags [F][N], projection matrices [F][12]
    double* d_vdmin = nullptr; int *d_vvis = nullptr, *d_vext = nullptr;
    // perception front-end workspace (tdlo_frontend.cuh)
    long long fe_cells_opt = 0, fe_cells_cap = 0, fe_tiles_cap = 0;
    int *d_fe_bbox = nullptr, *d_fe_dims = nullptr, *d_fe_tilecnt = nullptr, *d_fe_status = nullptr;
    long long *d_fe_cellbase = nullptr, *d_fe_tilebase = nullptr, *d_fe_acc = nullptr, *d_fe_tileoff = nullptr;
    unsigned char *d_fe_bgr = nullptr, *d_fe_occl = nullptr; unsigned short* d_fe_depth = nullptr; double* d_fe_proj = nullptr; long long fe_pix_cap = 0;
    // task-queue engine (tdlo_taskq.cuh)
    int tq_chunk = 0;               // raw points per chunk task (0 = automatic: 1024, or 2048 / 4096 for large batches)
    int tq_threads = 256;           // threads per CTA (2 CTAs/SM, 128 registers)
    int tq_inflight = 0;            // frames in flight (0 = automatic)
    double tq_zcut = 100.0;         // Gaussian truncation exponent (745.2 = exact zeros only)
    double tq_zrel = 45.0;          // relative truncation exponent (745.2 = off)
    int tq_solver = 0;              // M-step solve: 0 = automatic, 1 = dense always, 2 = structured always (TDLO_OPT_SOLVER)
    double watchdog_ms = 20000.0;   // a CTA waiting longer than this for a task aborts the launch (0 = off)
    cudaStream_t last_stream = nullptr; bool launched = false;
    double* d_fscratch = nullptr; long long fstride = 0;
    double *d_part = nullptr, *d_dminp = nullptr, *d_gath = nullptr; int* d_nkept = nullptr; double4* d_tsph = nullptr;
    unsigned long long* d_q = nullptr; unsigned qcap = 0; long long tq_chunks_cap = 0; int tq_alloc_chunk = 0;
    int32_t info[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    char err[512] = {0};
};

static char g_create_err[512] = {0};

static int fail(tdlo_ctx* ctx, int code, const char* fmt, ...) {
    char* dst = ctx ? ctx->err : g_create_err;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) return fail(ctx, TDLO_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

template <class T>
static cudaError_t dalloc(T** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(n, 1) * sizeof(T)); }

extern "C" const char* tdlo_version(void) { return "trackdlo_b200 0.1 (sm_100a)"; }

extern "C" const char* tdlo_last_error(const tdlo_ctx* ctx) { return ctx ? ctx->err : g_create_err; }

extern "C" void tdlo_destroy(tdlo_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    void* ptrs[] = {ctx->d_X, ctx->d_Y, ctx->d_sigma2, ctx->d_priors, ctx->d_H, ctx->d_W, ctx->d_rest, ctx->d_guide,
                    ctx->d_priors_out, ctx->d_packed, ctx->d_xoff, ctx->d_visoff, ctx->d_extoff, ctx->d_nnodes, ctx->d_npriors,
                    ctx->d_nvis, ctx->d_iters, ctx->d_status, ctx->d_vis, ctx->d_ext, ctx->d_npri_out, ctx->d_state,
                    ctx->d_Xc, ctx->d_bkt, ctx->d_exp_tab, ctx->d_prof_buf,
                    ctx->d_fscratch, ctx->d_part, ctx->d_dminp, ctx->d_gath, ctx->d_nkept, ctx->d_q, ctx->d_tsph,
                    ctx->d_vbits, ctx->d_vtmp, ctx->d_vcnt, ctx->d_vslice, ctx->d_vfree, ctx->d_vproj, ctx->d_vdmin, ctx->d_vvis, ctx->d_vext,
                    ctx->d_fe_bbox, ctx->d_fe_dims, ctx->d_fe_tilecnt, ctx->d_fe_status, ctx->d_fe_cellbase, ctx->d_fe_tilebase, ctx->d_fe_acc,
                    ctx->d_fe_tileoff, ctx->d_fe_bgr, ctx->d_fe_occl, ctx->d_fe_depth, ctx->d_fe_proj};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->ev_small) cudaEventDestroy(ctx->ev_small);
    if (ctx->ev_copy) cudaEventDestroy(ctx->ev_copy);
    if (ctx->d_ready) cudaFree(ctx->d_ready);
    if (ctx->h_ready_vals) cudaFreeHost(ctx->h_ready_vals);
    delete ctx;
}

extern "C" int tdlo_create(tdlo_ctx** out, int device, int32_t max_frames, int32_t max_nodes, int64_t max_points_total) {
    tdlo_ctx* ctx = nullptr;
    if (!out) return fail(nullptr, TDLO_ERR_INVALID, "tdlo_create: out is NULL");
    *out = nullptr;
    if (max_frames < 1 || max_nodes < 4 || max_nodes > TDLO_MAX_NODES || max_points_total < 1)
        return fail(nullptr, TDLO_ERR_INVALID, "tdlo_create: bad capacities (frames=%d nodes=%d points=%lld)", max_frames,
                    max_nodes, (long long)max_points_total);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, TDLO_ERR_CUDA, "tdlo_create: no CUDA device (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, TDLO_ERR_INVALID, "tdlo_create: device %d out of range", device);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return fail(nullptr, TDLO_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, TDLO_ERR_CUDA, "tdlo_create: device %d is sm_%d%d; this build targets sm_100a only", device,
                    prop.major, prop.minor);
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(nullptr, TDLO_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));

    ctx = new tdlo_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->max_frames = max_frames;
    ctx->max_nodes = max_nodes;
    ctx->max_points = max_points_total;
    const size_t F = max_frames, N = max_nodes, P = (size_t)max_points_total;
#define CKC(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e2_ = (call);                                                                   \
        if (e2_ != cudaSuccess) {                                                                   \
            fail(nullptr, e2_ == cudaErrorMemoryAllocation ? TDLO_ERR_NOMEM : TDLO_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e2_)); \
            tdlo_destroy(ctx);                                                                      \
            return e2_ == cudaErrorMemoryAllocation ? TDLO_ERR_NOMEM : TDLO_ERR_CUDA;              \
        }                                                                                           \
    } while (0)
    CKC(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CKC(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    CKC(cudaEventCreateWithFlags(&ctx->ev_small, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming));
    CKC(dalloc(&ctx->d_ready, 1));
    CKC(cudaMallocHost(reinterpret_cast<void**>(&ctx->h_ready_vals), 32 * sizeof(int)));     // [0..15] constants, [20] abort flag read-back
    for (int i = 0; i < 32; i++) ctx->h_ready_vals[i] = i < 16 ? i : 0;
    CKC(dalloc(&ctx->d_X, P * 3));
    CKC(dalloc(&ctx->d_Xc, P * 3));
    CKC(dalloc(&ctx->d_bkt, P));
    CKC(dalloc(&ctx->d_xoff, F + 1));
    CKC(dalloc(&ctx->d_Y, F * N * 3));
    CKC(dalloc(&ctx->d_sigma2, F));
    CKC(dalloc(&ctx->d_priors, F * N * 4));
    ctx->priors_cap = F * N * 4;
    CKC(dalloc(&ctx->d_W, F * N * 3));
    CKC(dalloc(&ctx->d_rest, F * N));
    CKC(dalloc(&ctx->d_guide, F * N * 3));
    CKC(dalloc(&ctx->d_priors_out, F * N * 8));
    CKC(dalloc(&ctx->d_visoff, F + 1));
    CKC(dalloc(&ctx->d_extoff, F + 1));
    CKC(dalloc(&ctx->d_nnodes, F));
    CKC(dalloc(&ctx->d_npriors, F));
    CKC(dalloc(&ctx->d_nvis, F));
    CKC(dalloc(&ctx->d_iters, F * 2));
    CKC(dalloc(&ctx->d_status, F));
    CKC(dalloc(&ctx->d_vis, F * N));
    CKC(dalloc(&ctx->d_ext, F * N));
    CKC(dalloc(&ctx->d_npri_out, F));
    CKC(dalloc(&ctx->d_state, F));
    CKC(dalloc(&ctx->d_prof_buf, 16));
    // exp table 2^(j/2048) (correctly rounded from long double)
    std::vector<double> tab(EXP_TAB);
    for (int j = 0; j < EXP_TAB; j++) tab[j] = (double)exp2l((long double)j / (long double)EXP_TAB);
    CKC(dalloc(&ctx->d_exp_tab, EXP_TAB));
    CKC(cudaMemcpy(ctx->d_exp_tab, tab.data(), sizeof(double) * EXP_TAB, cudaMemcpyHostToDevice));
#undef CKC
    *out = ctx;
    return TDLO_OK;
}

extern "C" int tdlo_last_launch_info(const tdlo_ctx* ctx, int32_t info[8]) {
    if (!ctx || !info) return TDLO_ERR_INVALID;
    memcpy(info, ctx->info, sizeof(ctx->info));
    return TDLO_OK;
}

// ---------------------------------------------------------------------------------------------
// launch
// ---------------------------------------------------------------------------------------------
// kernel variants, one translation unit each (tdlo_tq_inst_*.cu): <node passes, threads, resident CTAs>
namespace tdlo {
#define TDLO_TQ_DECL(NAME) cudaError_t NAME##_prepare(int smem, int* occ); cudaError_t NAME##_launch(int grid, int smem, cudaStream_t s, const TqArgs& t);
TDLO_TQ_DECL(tq_2_256_2) TDLO_TQ_DECL(tq_4_256_2) TDLO_TQ_DECL(tq_8_256_2)
#undef TDLO_TQ_DECL
}
struct TqVariant { int npass, threads; cudaError_t (*prepare)(int, int*); cudaError_t (*launch)(int, int, cudaStream_t, const TqArgs&); };

// Task-queue engine: one persistent launch, grid = SMs x resident CTAs, no clusters.
static int launch_tq(tdlo_ctx* ctx, KArgs& a, cudaStream_t stream) {
    CK(cudaSetDevice(ctx->device));
    if (a.n_frames > 0x1ffff) return fail(ctx, TDLO_ERR_INVALID, "task-queue engine: at most 131071 frames per call (got %d)", a.n_frames);
    const int nmax = a.nmax;
    const int threads = ctx->tq_threads;
    // ---- chunk size: small chunks shorten a frame's E-wave (few frames: latency bound), large chunks amortise the
    // per-task overhead (many frames: throughput bound).  Measured on C2 / C4 shapes: scripts/c4_probe.py.
    int chunk = ctx->tq_chunk;
    if (chunk == 0) {
        // chosen from the context CAPACITY (not the batch), so that host- and device-pointer entry points split the
        // points identically and stay bit-identical to each other
        const long long est = ctx->max_points;
        const long long slots = 2LL * ctx->sm_count;
        chunk = 1024;                                  // the largest chunk that still gives every CTA slot >= 2 tasks per wave;
        if (est / 2048 >= 2 * slots) chunk = 2048;     // small contexts (a single live sequence): short tasks, the E-wave
        if (est / 4096 >= 2 * slots) chunk = 4096;     // latency is on the frame's critical path
        if (est < 512 * slots) chunk = 512;
        if (est < 256 * slots) chunk = 256;
        // (a frame that fits ONE chunk -- at most 256 points: one 32-point tile per warp -- is run inline by the CTA that
        // owns it, without any queue traffic; larger single-CTA chunks lose: a tile is a ~9 k-cycle dependent chain for
        // its warp, so 2000 points on one CTA take 4x longer than spread over eight -- measured, profiles/r2_production_regime.jsonl)
    }
    // ---- workspace (allocated on first use / when the chunk size changes)
    if (!ctx->d_fscratch || ctx->tq_alloc_chunk != chunk) {
        void* old[] = {ctx->d_fscratch, ctx->d_part, ctx->d_dminp, ctx->d_gath, ctx->d_nkept, ctx->d_q, ctx->d_tsph};
        CK(cudaDeviceSynchronize());
        for (void* p : old) if (p) cudaFree(p);
        ctx->d_fscratch = nullptr; ctx->d_part = ctx->d_dminp = ctx->d_gath = nullptr; ctx->d_nkept = nullptr; ctx->d_q = nullptr; ctx->d_tsph = nullptr;
        const TqScr sc = tq_scr_layout(ctx->max_nodes);
        ctx->fstride = sc.total;
        ctx->tq_chunks_cap = ctx->max_points / chunk + ctx->max_frames + 2;
        unsigned cap = 1024;
        while ((long long)cap < 2 * (ctx->tq_chunks_cap + 4LL * ctx->sm_count + ctx->max_frames + 64)) cap <<= 1;
        ctx->qcap = cap;
#define CKA(call)                                                                                   \
        do { cudaError_t e2_ = (call); if (e2_ != cudaSuccess) return fail(ctx, e2_ == cudaErrorMemoryAllocation ? TDLO_ERR_NOMEM : TDLO_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e2_)); } while (0)
        CKA(dalloc(&ctx->d_fscratch, (size_t)ctx->fstride * ctx->max_frames));
        CKA(dalloc(&ctx->d_part, (size_t)ctx->tq_chunks_cap * (4 * ctx->max_nodes + 4)));
        CKA(dalloc(&ctx->d_dminp, (size_t)ctx->tq_chunks_cap * ctx->max_nodes));
        CKA(dalloc(&ctx->d_gath, (size_t)ctx->tq_chunks_cap));
        CKA(dalloc(&ctx->d_nkept, (size_t)ctx->tq_chunks_cap));
        CKA(dalloc(&ctx->d_q, (size_t)cap + 8));
        CKA(dalloc(&ctx->d_tsph, (size_t)ctx->tq_chunks_cap * (chunk / 32)));
#undef CKA
        ctx->tq_alloc_chunk = chunk;
    }
    // kernel variants: <node passes, threads, resident CTAs>; the shared-memory layout is sized for 32*passes nodes
    static const TqVariant kVariants[] = {{2, 256, tq_2_256_2_prepare, tq_2_256_2_launch},
                                          {4, 256, tq_4_256_2_prepare, tq_4_256_2_launch}, {8, 256, tq_8_256_2_prepare, tq_8_256_2_launch}};
    const TqVariant& kv = kVariants[nmax <= 64 ? 0 : (nmax <= 128 ? 1 : 2)];
    const int npass = kv.npass, threads_eff = kv.threads;
    const TqSmemL L = tq_smem_layout(32 * npass, threads_eff / 32);
    if (L.total > 227 * 1024) return fail(ctx, TDLO_ERR_INVALID, "node count %d does not fit shared memory", nmax);
    // development aid (scripts/occupancy_probe.py): TDLO_DEV_SMEM_PAD=<bytes> pads the dynamic shared memory to lower the
    // number of resident CTAs per SM
    static const int dev_pad = getenv("TDLO_DEV_SMEM_PAD") ? atoi(getenv("TDLO_DEV_SMEM_PAD")) : 0;
    const int smem_launch = std::min(L.total + dev_pad, 227 * 1024);
    int occ = 0;
    CK(kv.prepare(smem_launch, &occ));
    if (occ < 1) return fail(ctx, TDLO_ERR_CUDA, "task-queue kernel does not fit (smem %d B, %d threads)", L.total, threads_eff);
    const int grid = ctx->sm_count * occ;
    TqArgs t;
    memset(&t, 0, sizeof(t));
    a.Xc = ctx->d_Xc; a.bkt = ctx->d_bkt; a.scr_nodes = ctx->max_nodes; a.prof = ctx->d_prof;
    a.exp_tab = ctx->d_exp_tab; a.max_points = ctx->max_points;
    t.k = a;
    t.chunk = chunk;
    t.inflight = ctx->tq_inflight > 0 ? ctx->tq_inflight : std::max(grid / 2, 64);
    t.inflight = std::min(std::min(t.inflight, grid), a.n_frames);
    t.zcut = ctx->tq_zcut;
    t.zrel = ctx->tq_zrel;
    t.solver = ctx->tq_solver;
    t.qctl = ctx->d_q; t.qslots = ctx->d_q + 8; t.qmask = ctx->qcap - 1;
    t.fscratch = ctx->d_fscratch; t.fstride = ctx->fstride;
    t.part = ctx->d_part; t.dminp = ctx->d_dminp; t.gath = ctx->d_gath; t.nkept = ctx->d_nkept; t.tsph = ctx->d_tsph;
    t.part_stride = 4 * ctx->max_nodes + 4;
    t.ready = ctx->cur_ready; t.ready_frames = ctx->cur_ready_frames;
    t.L = L;
    t.watchdog_ns = (unsigned long long)(ctx->watchdog_ms * 1e6);
    CK(cudaMemsetAsync(ctx->d_q, 0, ((size_t)ctx->qcap + 8) * sizeof(unsigned long long), stream));
    ctx->last_stream = stream; ctx->launched = true;
    CK(kv.launch(grid, smem_launch, stream, t));
    ctx->info[0] = 1; ctx->info[1] = grid; ctx->info[2] = threads_eff; ctx->info[3] = L.total; ctx->info[4] = chunk;
    ctx->info[5] = 1; ctx->info[6] = occ; ctx->info[7] = ctx->sm_count;
    return TDLO_OK;
}

static int launch(tdlo_ctx* ctx, KArgs& a, cudaStream_t stream) { return launch_tq(ctx, a, stream); }

static CpdP to_dev(const tdlo_cpd_params& p) {
    CpdP d;
    d.beta = p.beta; d.lambda = p.lambda; d.gamma = p.lle_weight; d.mu = p.mu; d.tol = p.tol; d.alpha = p.alpha;
    d.k_vis = p.k_vis; d.tau = p.visibility_threshold; d.prune_radius = p.prune_radius;
    d.max_iter = p.max_iter; d.include_lle = p.include_lle;
    return d;
}

static int check_params(tdlo_ctx* ctx, double mu, double beta, int max_iter, double radius) {
    if (!(mu > 0.0 && mu < 1.0)) return fail(ctx, TDLO_ERR_INVALID, "mu must be in (0,1)");
    if (!(beta > 0.0)) return fail(ctx, TDLO_ERR_INVALID, "beta must be > 0");
    if (max_iter < 0) return fail(ctx, TDLO_ERR_INVALID, "max_iter must be >= 0");
    if (!(radius > 0.0)) return fail(ctx, TDLO_ERR_INVALID, "prune_radius must be > 0 (reference: 0.1)");
    return TDLO_OK;
}

extern "C" int tdlo_cpd_lle_batched_device(tdlo_ctx* ctx, const tdlo_cpd_batch* b, const tdlo_cpd_params* p, void* stream) {
    if (!ctx) return TDLO_ERR_INVALID;
    if (!b || !p) return fail(ctx, TDLO_ERR_INVALID, "null batch/params");
    if (b->n_frames < 0 || b->n_frames > ctx->max_frames) return fail(ctx, TDLO_ERR_INVALID, "n_frames %d exceeds capacity %d", b->n_frames, ctx->max_frames);
    if (b->node_stride < 4 || b->node_stride > ctx->max_nodes) return fail(ctx, TDLO_ERR_INVALID, "node_stride %d outside [4,%d]", b->node_stride, ctx->max_nodes);
    if (!b->X || !b->x_offsets || !b->Y || !b->sigma2) return fail(ctx, TDLO_ERR_INVALID, "X, x_offsets, Y, sigma2 are required");
    int rc = check_params(ctx, p->mu, p->beta, p->max_iter, p->prune_radius);
    if (rc) return rc;
    if (b->n_frames == 0) { ctx->info[5] = 0; return TDLO_OK; }
    KArgs a;
    memset(&a, 0, sizeof(a));
    a.mode = 0;
    a.n_frames = b->n_frames; a.node_stride = b->node_stride; a.nmax = b->node_stride;
    a.X = b->X; a.x_off = reinterpret_cast<const long long*>(b->x_offsets);
    a.n_nodes = b->n_nodes; a.Y = b->Y; a.sigma2 = b->sigma2;
    a.priors = b->priors; a.n_priors = b->n_priors; a.n_visible = b->n_visible; a.H = b->H;
    a.priors_stride = b->priors_stride > 0 ? b->priors_stride : b->node_stride;
    if (b->priors_stride < 0) return fail(ctx, TDLO_ERR_INVALID, "priors_stride must be >= 0");
    a.W = b->W; a.iters = b->iters; a.status = b->status;
    a.p0 = to_dev(*p);
    a.p1 = a.p0;
    return launch(ctx, a, (cudaStream_t)stream);
}

extern "C" int tdlo_tracking_step_batched_device(tdlo_ctx* ctx, const tdlo_track_batch* b, const tdlo_track_params* p, void* stream) {
    if (!ctx) return TDLO_ERR_INVALID;
    if (!b || !p) return fail(ctx, TDLO_ERR_INVALID, "null batch/params");
    if (b->n_frames < 0 || b->n_frames > ctx->max_frames) return fail(ctx, TDLO_ERR_INVALID, "n_frames %d exceeds capacity %d", b->n_frames, ctx->max_frames);
    if (b->n_nodes < 4 || b->n_nodes > ctx->max_nodes) return fail(ctx, TDLO_ERR_INVALID, "n_nodes %d outside [4,%d]", b->n_nodes, ctx->max_nodes);
    if (!b->X || !b->x_offsets || !b->Y || !b->sigma2 || !b->geodesic_coord || !b->visible || !b->visible_offsets ||
        !b->visible_ext || !b->visible_ext_offsets)
        return fail(ctx, TDLO_ERR_INVALID, "X, x_offsets, Y, sigma2, geodesic_coord and both visibility lists are required");
    int rc = check_params(ctx, p->mu, p->beta, p->max_iter, p->prune_radius);
    if (rc) return rc;
    if (!(p->beta_pre_proc > 0.0)) return fail(ctx, TDLO_ERR_INVALID, "beta_pre_proc must be > 0");
    if (b->n_frames == 0) { ctx->info[5] = 0; return TDLO_OK; }
    KArgs a;
    memset(&a, 0, sizeof(a));
    a.mode = 1;
    a.n_frames = b->n_frames; a.node_stride = b->n_nodes; a.nmax = b->n_nodes;
    a.X = b->X; a.x_off = reinterpret_cast<const long long*>(b->x_offsets);
    a.Y = b->Y; a.sigma2 = b->sigma2; a.H = b->H_pre;
    a.iters = b->iters; a.status = b->status;
    a.rest = b->geodesic_coord;
    a.vis = b->visible; a.vis_off = reinterpret_cast<const long long*>(b->visible_offsets);
    a.ext = b->visible_ext; a.ext_off = reinterpret_cast<const long long*>(b->visible_ext_offsets);
    a.guide_out = b->guide_nodes; a.priors_out = b->priors; a.n_priors_out = b->n_priors; a.state_out = b->state;
    a.packed_out = b->packed_results;
    // pre-processing call: cpd_lle(X, guide, s2, beta_pre, lambda_pre, lle_weight, mu, max_iter, tol, true)
    // with the header defaults alpha=0, k_vis=0, visibility_threshold=0.01 (trackdlo.cpp:927, trackdlo.h:91-95)
    a.p0.beta = p->beta_pre_proc; a.p0.lambda = p->lambda_pre_proc; a.p0.gamma = p->lle_weight; a.p0.mu = p->mu;
    a.p0.tol = p->tol; a.p0.alpha = 0.0; a.p0.k_vis = 0.0; a.p0.tau = 0.01; a.p0.prune_radius = p->prune_radius;
    a.p0.max_iter = p->max_iter; a.p0.include_lle = 1;
    // main call (trackdlo.cpp:998)
    a.p1.beta = p->beta; a.p1.lambda = p->lambda; a.p1.gamma = p->lle_weight; a.p1.mu = p->mu; a.p1.tol = p->tol;
    a.p1.alpha = p->alpha; a.p1.k_vis = p->k_vis; a.p1.tau = p->visibility_threshold; a.p1.prune_radius = p->prune_radius;
    a.p1.max_iter = p->max_iter; a.p1.include_lle = 0;
    return launch(ctx, a, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// host-pointer entry points: copy in, run, copy out, synchronise
// ---------------------------------------------------------------------------------------------
#define H2D(dst, src, bytes) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream))
#define D2H(dst, src, bytes) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream))

// The persistent kernel's watchdog (tdlo_taskq.cuh: tq_watchdog) leaves a flag behind instead of hanging the caller:
// queue its read-back behind the kernel, synchronise, and turn it into an error.
static int sync_and_check(tdlo_ctx* ctx, cudaStream_t stream) {
    int* h_abort = ctx->h_ready_vals + 20;
    if (ctx->launched && ctx->d_q) CK(cudaMemcpyAsync(h_abort, reinterpret_cast<int*>(ctx->d_q + 3) + 1, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (ctx->launched && *h_abort) {
        *h_abort = 0;
        return fail(ctx, TDLO_ERR_CUDA, "watchdog: the task queue made no progress for %.0f ms (lost task or missing upload); results of this call are invalid", ctx->watchdog_ms);
    }
    return TDLO_OK;
}

// ---------------------------------------------------------------------------------------------
// multi-rank gather of the packed records over the caller's NCCL communicator (NCCL is loaded lazily: no link dependency)
// ---------------------------------------------------------------------------------------------
extern "C" int tdlo_all_gather_packed(tdlo_ctx* ctx, void* nccl_comm, double* packed_all, int32_t frames_per_rank, int32_t n_nodes,
                                      int32_t rank, void* stream) {
    if (!ctx) return TDLO_ERR_INVALID;
    if (!nccl_comm || !packed_all || frames_per_rank < 0 || n_nodes < 1 || rank < 0) return fail(ctx, TDLO_ERR_INVALID, "tdlo_all_gather_packed: bad arguments");
    // ncclResult_t ncclAllGather(const void* sendbuff, void* recvbuff, size_t sendcount, ncclDataType_t datatype, ncclComm_t comm, cudaStream_t stream)
    typedef int (*all_gather_fn)(const void*, void*, size_t, int, void*, cudaStream_t);
    typedef const char* (*errstr_fn)(int);
    static all_gather_fn p_all_gather = nullptr;
    static errstr_fn p_errstr = nullptr;
    if (!p_all_gather) {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return fail(ctx, TDLO_ERR_CUDA, "tdlo_all_gather_packed: cannot load libnccl.so.2 (%s)", dlerror());
        p_all_gather = reinterpret_cast<all_gather_fn>(dlsym(h, "ncclAllGather"));
        p_errstr = reinterpret_cast<errstr_fn>(dlsym(h, "ncclGetErrorString"));
        if (!p_all_gather) return fail(ctx, TDLO_ERR_CUDA, "tdlo_all_gather_packed: ncclAllGather not found in libnccl");
    }
    CK(cudaSetDevice(ctx->device));
    const size_t cnt = (size_t)frames_per_rank * (3 * (size_t)n_nodes + 4);
    constexpr int kNcclFloat64 = 8;                     // ncclFloat64 / ncclDouble (nccl.h)
    const int rc = p_all_gather(packed_all + (size_t)rank * cnt, packed_all, cnt, kNcclFloat64, nccl_comm, (cudaStream_t)stream);
    if (rc != 0) return fail(ctx, TDLO_ERR_CUDA, "ncclAllGather: %s", p_errstr ? p_errstr(rc) : "error");
    return TDLO_OK;
}

extern "C" int tdlo_synchronize(tdlo_ctx* ctx) {
    if (!ctx) return TDLO_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    return sync_and_check(ctx, ctx->last_stream);
}

// Uploads the concatenated point clouds.  With the task-queue engine and a batch worth splitting, the upload runs in
// (up to) 8 groups of frames on the copy stream, each followed by a progress flag; the persistent kernel starts right
// away on the main stream and a frame's first task waits for its group's flag, so that the transfer of the later
// frames overlaps the registration of the earlier ones.  Call after the small arrays are queued on ctx->stream.
static int upload_points(tdlo_ctx* ctx, const double* X, const int64_t* x_offsets, int F) {
    const long long total = x_offsets[F];
    ctx->cur_ready = nullptr; ctx->cur_ready_frames = 0;
    const int G = (F >= 16 && total >= 200000) ? 8 : 1;
    if (G == 1) {
        if (total > 0) H2D(ctx->d_X, X, (size_t)total * 3 * sizeof(double));
        return TDLO_OK;
    }
    const int per = (F + G - 1) / G;
    CK(cudaMemsetAsync(ctx->d_ready, 0, sizeof(int), ctx->stream));
    CK(cudaEventRecord(ctx->ev_small, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_small, 0));
    for (int g = 0; g * per < F; g++) {
        const int f0 = g * per, f1 = std::min(F, f0 + per);
        const long long p0 = x_offsets[f0], p1 = x_offsets[f1];
        if (p1 > p0) CK(cudaMemcpyAsync(ctx->d_X + p0 * 3, X + p0 * 3, (size_t)(p1 - p0) * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->copy_stream));
        CK(cudaMemcpyAsync(ctx->d_ready, ctx->h_ready_vals + g + 1, sizeof(int), cudaMemcpyHostToDevice, ctx->copy_stream));
    }
    CK(cudaEventRecord(ctx->ev_copy, ctx->copy_stream));
    ctx->cur_ready = ctx->d_ready; ctx->cur_ready_frames = per;
    return TDLO_OK;
}
// after the kernel has been queued: later work on the main stream must not overtake the copy stream
static int upload_points_join(tdlo_ctx* ctx) {
    if (ctx->cur_ready) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy, 0));
    ctx->cur_ready = nullptr; ctx->cur_ready_frames = 0;
    return TDLO_OK;
}

extern "C" int tdlo_cpd_lle_batched(tdlo_ctx* ctx, const tdlo_cpd_batch* b, const tdlo_cpd_params* p) {
    if (!ctx) return TDLO_ERR_INVALID;
    if (!b || !p) return fail(ctx, TDLO_ERR_INVALID, "null batch/params");
    if (b->n_frames < 0 || b->n_frames > ctx->max_frames) return fail(ctx, TDLO_ERR_INVALID, "n_frames %d exceeds capacity %d", b->n_frames, ctx->max_frames);
    if (b->n_frames == 0) return TDLO_OK;
    if (b->node_stride < 4 || b->node_stride > ctx->max_nodes) return fail(ctx, TDLO_ERR_INVALID, "node_stride %d outside [4,%d]", b->node_stride, ctx->max_nodes);
    if (!b->X || !b->x_offsets || !b->Y || !b->sigma2) return fail(ctx, TDLO_ERR_INVALID, "X, x_offsets, Y, sigma2 are required");
    { int rcp = check_params(ctx, p->mu, p->beta, p->max_iter, p->prune_radius); if (rcp) return rcp; }   // before any copy is queued
    const size_t F = b->n_frames, S = b->node_stride;
    const size_t PS = b->priors_stride > 0 ? (size_t)b->priors_stride : S;
    if (b->priors_stride < 0 || PS > 4u * TDLO_MAX_NODES) return fail(ctx, TDLO_ERR_INVALID, "priors_stride %d outside [0,%d]", b->priors_stride, 4 * TDLO_MAX_NODES);
    if (b->priors && b->n_priors) for (size_t f = 0; f < F; f++) if (b->n_priors[f] < 0 || (size_t)b->n_priors[f] > PS)
        return fail(ctx, TDLO_ERR_INVALID, "n_priors[%zu] = %d outside [0, priors_stride = %zu]", f, b->n_priors[f], PS);
    const long long total = b->x_offsets[F];
    if (b->x_offsets[0] != 0) return fail(ctx, TDLO_ERR_INVALID, "x_offsets[0] must be 0");
    for (size_t f = 0; f < F; f++) if (b->x_offsets[f + 1] < b->x_offsets[f]) return fail(ctx, TDLO_ERR_INVALID, "x_offsets not monotone");
    if (total > ctx->max_points) return fail(ctx, TDLO_ERR_INVALID, "%lld points exceed capacity %lld", total, ctx->max_points);
    if (b->n_nodes) for (size_t f = 0; f < F; f++) if (b->n_nodes[f] > (int)S || b->n_nodes[f] < 0) return fail(ctx, TDLO_ERR_INVALID, "n_nodes[%zu] outside [0,node_stride]", f);
    CK(cudaSetDevice(ctx->device));
    if (b->H && !ctx->d_H) CK(dalloc(&ctx->d_H, (size_t)ctx->max_frames * ctx->max_nodes * ctx->max_nodes));
    H2D(ctx->d_xoff, b->x_offsets, (F + 1) * sizeof(long long));
    H2D(ctx->d_Y, b->Y, F * S * 3 * sizeof(double));
    H2D(ctx->d_sigma2, b->sigma2, F * sizeof(double));
    if (b->n_nodes) H2D(ctx->d_nnodes, b->n_nodes, F * sizeof(int));
    if (b->priors) {
        if (PS * 4 * (size_t)ctx->max_frames > ctx->priors_cap) {     // longer prior lists than one row per node: grow the staging buffer
            CK(cudaStreamSynchronize(ctx->stream));
            if (ctx->d_priors) cudaFree(ctx->d_priors);
            ctx->d_priors = nullptr; ctx->priors_cap = 0;
            CK(dalloc(&ctx->d_priors, PS * 4 * (size_t)ctx->max_frames));
            ctx->priors_cap = PS * 4 * (size_t)ctx->max_frames;
        }
        H2D(ctx->d_priors, b->priors, F * PS * 4 * sizeof(double));
    }
    if (b->n_priors) H2D(ctx->d_npriors, b->n_priors, F * sizeof(int));
    if (b->n_visible) H2D(ctx->d_nvis, b->n_visible, F * sizeof(int));
    if (b->H) H2D(ctx->d_H, b->H, F * S * S * sizeof(double));
    { int rcu = upload_points(ctx, b->X, b->x_offsets, (int)F); if (rcu) return rcu; }
    tdlo_cpd_batch d = *b;
    d.X = ctx->d_X; d.x_offsets = reinterpret_cast<const int64_t*>(ctx->d_xoff); d.Y = ctx->d_Y; d.sigma2 = ctx->d_sigma2;
    d.n_nodes = b->n_nodes ? ctx->d_nnodes : nullptr;
    d.priors = b->priors ? ctx->d_priors : nullptr;
    d.n_priors = b->n_priors ? ctx->d_npriors : nullptr;
    d.n_visible = b->n_visible ? ctx->d_nvis : nullptr;
    d.H = b->H ? ctx->d_H : nullptr;
    d.W = ctx->d_W; d.iters = ctx->d_iters; d.status = ctx->d_status;
    int rc = tdlo_cpd_lle_batched_device(ctx, &d, p, ctx->stream);
    { int rcj = upload_points_join(ctx); if (!rc) rc = rcj; }
    if (rc) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamSynchronize(ctx->stream); return rc; }   // no copy from the caller's buffers left in flight
    D2H(b->Y, ctx->d_Y, F * S * 3 * sizeof(double));
    D2H(b->sigma2, ctx->d_sigma2, F * sizeof(double));
    if (b->W) D2H(b->W, ctx->d_W, F * S * 3 * sizeof(double));
    if (b->iters) D2H(b->iters, ctx->d_iters, F * sizeof(int));
    if (b->status) D2H(b->status, ctx->d_status, F * sizeof(int));
    return sync_and_check(ctx, ctx->stream);
}

extern "C" int tdlo_tracking_step_batched(tdlo_ctx* ctx, const tdlo_track_batch* b, const tdlo_track_params* p) {
    if (!ctx) return TDLO_ERR_INVALID;
    if (!b || !p) return fail(ctx, TDLO_ERR_INVALID, "null batch/params");
    if (b->n_frames < 0 || b->n_frames > ctx->max_frames) return fail(ctx, TDLO_ERR_INVALID, "n_frames %d exceeds capacity %d", b->n_frames, ctx->max_frames);
    if (b->n_frames == 0) return TDLO_OK;
    if (b->n_nodes < 4 || b->n_nodes > ctx->max_nodes) return fail(ctx, TDLO_ERR_INVALID, "n_nodes %d outside [4,%d]", b->n_nodes, ctx->max_nodes);
    if (!b->X || !b->x_offsets || !b->Y || !b->sigma2 || !b->geodesic_coord || !b->visible || !b->visible_offsets ||
        !b->visible_ext || !b->visible_ext_offsets)
        return fail(ctx, TDLO_ERR_INVALID, "X, x_offsets, Y, sigma2, geodesic_coord and both visibility lists are required");
    { int rcp = check_params(ctx, p->mu, p->beta, p->max_iter, p->prune_radius); if (rcp) return rcp; }   // before any copy is queued
    if (!(p->beta_pre_proc > 0.0)) return fail(ctx, TDLO_ERR_INVALID, "beta_pre_proc must be > 0");
    const size_t F = b->n_frames, N = b->n_nodes;
    const long long total = b->x_offsets[F];
    if (b->x_offsets[0] != 0) return fail(ctx, TDLO_ERR_INVALID, "x_offsets[0] must be 0");
    for (size_t f = 0; f < F; f++) if (b->x_offsets[f + 1] < b->x_offsets[f]) return fail(ctx, TDLO_ERR_INVALID, "x_offsets not monotone");
    if (total > ctx->max_points) return fail(ctx, TDLO_ERR_INVALID, "%lld points exceed capacity %lld", total, ctx->max_points);
    for (size_t f = 0; f < F; f++) {
        const long long nv = b->visible_offsets[f + 1] - b->visible_offsets[f], ne = b->visible_ext_offsets[f + 1] - b->visible_ext_offsets[f];
        if (nv < 0 || nv > (long long)N || ne < 0 || ne > (long long)N) return fail(ctx, TDLO_ERR_INVALID, "frame %zu: visibility list longer than n_nodes", f);
        for (long long i = 0; i < nv; i++) { const int v = b->visible[b->visible_offsets[f] + i]; if (v < 0 || v >= (int)N) return fail(ctx, TDLO_ERR_INVALID, "frame %zu: visible node %d out of range", f, v); }
        for (long long i = 0; i < ne; i++) {
            const int v = b->visible_ext[b->visible_ext_offsets[f] + i];
            if (v < 0 || v >= (int)N) return fail(ctx, TDLO_ERR_INVALID, "frame %zu: visible_ext node %d out of range", f, v);
            if (i > 0 && v <= b->visible_ext[b->visible_ext_offsets[f] + i - 1]) return fail(ctx, TDLO_ERR_INVALID, "frame %zu: visible_ext must be strictly ascending", f);
        }
    }
    CK(cudaSetDevice(ctx->device));
    if (b->H_pre && !ctx->d_H) CK(dalloc(&ctx->d_H, (size_t)ctx->max_frames * ctx->max_nodes * ctx->max_nodes));
    H2D(ctx->d_xoff, b->x_offsets, (F + 1) * sizeof(long long));
    H2D(ctx->d_Y, b->Y, F * N * 3 * sizeof(double));
    H2D(ctx->d_sigma2, b->sigma2, F * sizeof(double));
    H2D(ctx->d_rest, b->geodesic_coord, F * N * sizeof(double));
    H2D(ctx->d_visoff, b->visible_offsets, (F + 1) * sizeof(long long));
    H2D(ctx->d_extoff, b->visible_ext_offsets, (F + 1) * sizeof(long long));
    if (b->visible_offsets[F] > 0) H2D(ctx->d_vis, b->visible, (size_t)b->visible_offsets[F] * sizeof(int));
    if (b->visible_ext_offsets[F] > 0) H2D(ctx->d_ext, b->visible_ext, (size_t)b->visible_ext_offsets[F] * sizeof(int));
    if (b->H_pre) H2D(ctx->d_H, b->H_pre, F * N * N * sizeof(double));
    { int rcu = upload_points(ctx, b->X, b->x_offsets, (int)F); if (rcu) return rcu; }
    tdlo_track_batch d = *b;
    d.X = ctx->d_X; d.x_offsets = reinterpret_cast<const int64_t*>(ctx->d_xoff); d.Y = ctx->d_Y; d.sigma2 = ctx->d_sigma2;
    d.geodesic_coord = ctx->d_rest;
    d.visible = ctx->d_vis; d.visible_offsets = reinterpret_cast<const int64_t*>(ctx->d_visoff);
    d.visible_ext = ctx->d_ext; d.visible_ext_offsets = reinterpret_cast<const int64_t*>(ctx->d_extoff);
    d.H_pre = b->H_pre ? ctx->d_H : nullptr;
    d.guide_nodes = ctx->d_guide; d.priors = ctx->d_priors_out; d.n_priors = ctx->d_npri_out;
    d.iters = ctx->d_iters; d.status = ctx->d_status; d.state = ctx->d_state;
    d.packed_results = nullptr;
    if (b->packed_results) {
        if (!ctx->d_packed) CK(dalloc(&ctx->d_packed, (size_t)ctx->max_frames * (3 * (size_t)ctx->max_nodes + 4)));
        d.packed_results = ctx->d_packed;
    }
    int rc = tdlo_tracking_step_batched_device(ctx, &d, p, ctx->stream);
    { int rcj = upload_points_join(ctx); if (!rc) rc = rcj; }
    if (rc) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamSynchronize(ctx->stream); return rc; }   // no copy from the caller's buffers left in flight
    D2H(b->Y, ctx->d_Y, F * N * 3 * sizeof(double));
    D2H(b->sigma2, ctx->d_sigma2, F * sizeof(double));
    if (b->guide_nodes) D2H(b->guide_nodes, ctx->d_guide, F * N * 3 * sizeof(double));
    if (b->priors) D2H(b->priors, ctx->d_priors_out, F * N * 8 * sizeof(double));
    if (b->n_priors) D2H(b->n_priors, ctx->d_npri_out, F * sizeof(int));
    if (b->iters) D2H(b->iters, ctx->d_iters, F * 2 * sizeof(int));
    if (b->status) D2H(b->status, ctx->d_status, F * sizeof(int));
    if (b->state) D2H(b->state, ctx->d_state, F * sizeof(int));
    if (b->packed_results) D2H(b->packed_results, ctx->d_packed, F * (3 * N + 4) * sizeof(double));
    return sync_and_check(ctx, ctx->stream);
}

// ---------------------------------------------------------------------------------------------
// visibility front-end (tdlo_visibility.cuh)
// ---------------------------------------------------------------------------------------------
static int vis_workspace(tdlo_ctx* ctx) {
    if (ctx->d_vbits) return TDLO_OK;
    const size_t F = ctx->max_frames, N = ctx->max_nodes;
    CK(dalloc(&ctx->d_vbits, F * N));
    CK(dalloc(&ctx->d_vtmp, 2 * F * N));
    CK(dalloc(&ctx->d_vcnt, 2 * F));
    CK(dalloc(&ctx->d_vslice, F + 1));
    CK(dalloc(&ctx->d_vfree, F * N));
    CK(dalloc(&ctx->d_vproj, F * 12));
    return TDLO_OK;
}

// Stream-ordered: the slice table is built on the device (no read-back of the offsets).
static int vis_launch(tdlo_ctx* ctx, const tdlo_vis_batch* b, cudaStream_t stream) {
    int rc = vis_workspace(ctx);
    if (rc) return rc;
    const int F = b->n_frames, N = b->n_nodes;
    VisArgs a;
    a.n_frames = F; a.n_nodes = N;
    a.X = b->X; a.x_off = reinterpret_cast<const long long*>(b->x_offsets); a.Y = b->Y; a.node_coord = b->node_coord;
    a.tau = b->visibility_threshold; a.d_vis = b->d_vis;
    a.dmin2_bits = ctx->d_vbits; a.tmp_vis = ctx->d_vtmp; a.tmp_ext = ctx->d_vtmp + (size_t)ctx->max_frames * ctx->max_nodes;
    a.counts = ctx->d_vcnt; a.dmin_out = b->dmin;
    a.slice_start = ctx->d_vslice; a.max_points = ctx->max_points;
    a.vis = b->visible; a.vis_off = reinterpret_cast<long long*>(b->visible_offsets);
    a.ext = b->visible_ext; a.ext_off = reinterpret_cast<long long*>(b->visible_ext_offsets);
    a.proj = b->proj; a.rows = b->rows; a.cols = b->cols; a.pixel_width = b->pixel_width;
    a.free_flag = b->proj ? (b->not_self_occluded ? b->not_self_occluded : ctx->d_vfree) : nullptr;
    CK(cudaMemsetAsync(ctx->d_vbits, 0x7f, (size_t)F * N * sizeof(unsigned long long), stream));   // 0x7f7f... = 1.4e306
    const long long max_slices = ctx->max_points / VIS_SLICE + F;
    const int grid = (int)std::max<long long>(1, std::min<long long>(max_slices, (long long)ctx->sm_count * 8));
    tdlo_vis_slices_kernel<<<1, 256, 0, stream>>>(a);
    tdlo_vis_dmin_kernel<<<grid, 256, 0, stream>>>(a);
    if (a.proj) tdlo_vis_selfocc_kernel<<<F, 256, 0, stream>>>(a);
    tdlo_vis_lists_kernel<<<(F + 63) / 64, 64, 0, stream>>>(a);
    tdlo_vis_compact_kernel<<<1, 256, 0, stream>>>(a);
    CK(cudaGetLastError());
    ctx->info[5] = 4;
    return TDLO_OK;
}

static int vis_check(tdlo_ctx* ctx, const tdlo_vis_batch* b) {
    if (!b) return fail(ctx, TDLO_ERR_INVALID, "null batch");
    if (b->n_frames < 0 || b->n_frames > ctx->max_frames) return fail(ctx, TDLO_ERR_INVALID, "n_frames %d exceeds capacity %d", b->n_frames, ctx->max_frames);
    if (b->n_nodes < 1 || b->n_nodes > ctx->max_nodes) return fail(ctx, TDLO_ERR_INVALID, "n_nodes %d outside [1,%d]", b->n_nodes, ctx->max_nodes);
    if (!b->X || !b->x_offsets || !b->Y || !b->node_coord || !b->visible || !b->visible_offsets || !b->visible_ext || !b->visible_ext_offsets)
        return fail(ctx, TDLO_ERR_INVALID, "X, x_offsets, Y, node_coord and the four output arrays are required");
    if (b->proj) {
        if (b->rows < 1 || b->cols < 1 || b->rows > (1 << 20) || b->cols > (1 << 20)) return fail(ctx, TDLO_ERR_INVALID, "self-occlusion test: bad image size %d x %d", b->rows, b->cols);
        if (b->pixel_width < 2 || b->pixel_width > 2 * SO_MAX_RADIUS) return fail(ctx, TDLO_ERR_INVALID, "self-occlusion test: pixel_width %d outside [2, %d]", b->pixel_width, 2 * SO_MAX_RADIUS);
    }
    return TDLO_OK;
}

extern "C" int tdlo_visibility_batched_device(tdlo_ctx* ctx, const tdlo_vis_batch* b, void* stream) {
    if (!ctx) return TDLO_ERR_INVALID;
    int rc = vis_check(ctx, b);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    if (b->n_frames == 0) return TDLO_OK;
    return vis_launch(ctx, b, (cudaStream_t)stream);
}

extern "C" int tdlo_visibility_batched(tdlo_ctx* ctx, const tdlo_vis_batch* b) {
    if (!ctx) return TDLO_ERR_INVALID;
    int rc = vis_check(ctx, b);
    if (rc) return rc;
    if (b->n_frames == 0) return TDLO_OK;
    const size_t F = b->n_frames, N = b->n_nodes;
    const long long total = b->x_offsets[F];
    if (b->x_offsets[0] != 0) return fail(ctx, TDLO_ERR_INVALID, "x_offsets[0] must be 0");
    for (size_t f = 0; f < F; f++) if (b->x_offsets[f + 1] < b->x_offsets[f]) return fail(ctx, TDLO_ERR_INVALID, "x_offsets not monotone");
    if (total > ctx->max_points) return fail(ctx, TDLO_ERR_INVALID, "%lld points exceed capacity %lld", total, ctx->max_points);
    CK(cudaSetDevice(ctx->device));
    if (!ctx->d_vdmin) {
        const size_t FM = ctx->max_frames, NM = ctx->max_nodes;
        CK(dalloc(&ctx->d_vdmin, FM * NM)); CK(dalloc(&ctx->d_vvis, FM * NM)); CK(dalloc(&ctx->d_vext, FM * NM));
    }
    H2D(ctx->d_X, b->X, (size_t)total * 3 * sizeof(double));
    H2D(ctx->d_xoff, b->x_offsets, (F + 1) * sizeof(long long));
    H2D(ctx->d_Y, b->Y, F * N * 3 * sizeof(double));
    H2D(ctx->d_rest, b->node_coord, F * N * sizeof(double));
    rc = vis_workspace(ctx);
    if (rc) return rc;
    if (b->proj) H2D(ctx->d_vproj, b->proj, F * 12 * sizeof(double));
    tdlo_vis_batch d = *b;
    if (b->proj) { d.proj = ctx->d_vproj; d.not_self_occluded = ctx->d_vfree; }
    d.X = ctx->d_X; d.x_offsets = reinterpret_cast<const int64_t*>(ctx->d_xoff); d.Y = ctx->d_Y; d.node_coord = ctx->d_rest;
    d.dmin = ctx->d_vdmin; d.visible = ctx->d_vvis; d.visible_offsets = reinterpret_cast<int64_t*>(ctx->d_visoff);
    d.visible_ext = ctx->d_vext; d.visible_ext_offsets = reinterpret_cast<int64_t*>(ctx->d_extoff);
    rc = vis_launch(ctx, &d, ctx->stream);
    if (rc) return rc;
    if (b->dmin) D2H(b->dmin, ctx->d_vdmin, F * N * sizeof(double));
    if (b->proj && b->not_self_occluded) D2H(b->not_self_occluded, ctx->d_vfree, F * N * sizeof(int));
    D2H(b->visible_offsets, ctx->d_visoff, (F + 1) * sizeof(long long));
    D2H(b->visible_ext_offsets, ctx->d_extoff, (F + 1) * sizeof(long long));
    D2H(b->visible, ctx->d_vvis, F * N * sizeof(int));
    D2H(b->visible_ext, ctx->d_vext, F * N * sizeof(int));
    CK(cudaStreamSynchronize(ctx->stream));
    return TDLO_OK;
}

// ---------------------------------------------------------------------------------------------
// perception front-end (tdlo_frontend.cuh)
// ---------------------------------------------------------------------------------------------
static int fe_workspace(tdlo_ctx* ctx) {
    long long want = ctx->fe_cells_opt;
    if (want <= 0) want = std::min<long long>(1LL << 25, std::max<long long>(1LL << 22, (long long)ctx->max_frames << 19));
    if (ctx->d_fe_acc && ctx->fe_cells_cap == want) return TDLO_OK;
    CK(cudaDeviceSynchronize());
    void* old[] = {ctx->d_fe_bbox, ctx->d_fe_dims, ctx->d_fe_tilecnt, ctx->d_fe_status, ctx->d_fe_cellbase, ctx->d_fe_tilebase, ctx->d_fe_acc, ctx->d_fe_tileoff};
    for (void* p : old) if (p) cudaFree(p);
    ctx->d_fe_bbox = ctx->d_fe_dims = ctx->d_fe_tilecnt = ctx->d_fe_status = nullptr;
    ctx->d_fe_cellbase = ctx->d_fe_tilebase = ctx->d_fe_acc = ctx->d_fe_tileoff = nullptr;
    const size_t F = ctx->max_frames;
    ctx->fe_cells_cap = want;
    ctx->fe_tiles_cap = want / FE_TILE + (long long)F + 1;
    CK(dalloc(&ctx->d_fe_bbox, F * 6)); CK(dalloc(&ctx->d_fe_dims, F * 4)); CK(dalloc(&ctx->d_fe_status, F));
    CK(dalloc(&ctx->d_fe_cellbase, F + 1)); CK(dalloc(&ctx->d_fe_tilebase, F + 1));
    CK(dalloc(&ctx->d_fe_acc, (size_t)want * 4));
    CK(dalloc(&ctx->d_fe_tilecnt, (size_t)ctx->fe_tiles_cap)); CK(dalloc(&ctx->d_fe_tileoff, (size_t)ctx->fe_tiles_cap));
    return TDLO_OK;
}

static int fe_check(tdlo_ctx* ctx, const tdlo_frontend_batch* b) {
    if (!b) return fail(ctx, TDLO_ERR_INVALID, "null batch");
    if (b->n_frames < 0 || b->n_frames > ctx->max_frames) return fail(ctx, TDLO_ERR_INVALID, "n_frames %d exceeds capacity %d", b->n_frames, ctx->max_frames);
    if (b->rows < 1 || b->cols < 1 || (long long)b->rows * b->cols > (1LL << 30)) return fail(ctx, TDLO_ERR_INVALID, "bad image size %d x %d", b->rows, b->cols);
    if (!b->bgr || !b->depth || !b->proj || !b->X || !b->x_offsets) return fail(ctx, TDLO_ERR_INVALID, "bgr, depth, proj, X, x_offsets are required");
    if (!(b->leaf_size > 0.0)) return fail(ctx, TDLO_ERR_INVALID, "leaf_size must be > 0 (reference: 0.008)");
    if (b->x_capacity < 0) return fail(ctx, TDLO_ERR_INVALID, "x_capacity must be >= 0");
    return TDLO_OK;
}

extern "C" int tdlo_point_cloud_batched_device(tdlo_ctx* ctx, const tdlo_frontend_batch* b, void* stream_) {
    if (!ctx) return TDLO_ERR_INVALID;
    int rc = fe_check(ctx, b);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    if (b->n_frames == 0) return TDLO_OK;
    rc = fe_workspace(ctx);
    if (rc) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    FeArgs a;
    memset(&a, 0, sizeof(a));
    a.n_frames = b->n_frames; a.rows = b->rows; a.cols = b->cols; a.multi = b->multi_color;
    a.bgr = b->bgr; a.depth = b->depth; a.occl = b->occlusion_bgr; a.proj = b->proj;
    for (int i = 0; i < 3; i++) { a.lo[i] = b->hsv_lower[i]; a.hi[i] = b->hsv_upper[i]; }
    a.inv_leaf = 1.0f / (float)b->leaf_size;                 // PCL: inverse_leaf_size_ = 1 / leaf_size_ in float
    a.bbox = ctx->d_fe_bbox; a.dims = ctx->d_fe_dims; a.cell_base = ctx->d_fe_cellbase; a.tile_base = ctx->d_fe_tilebase;
    a.acc = ctx->d_fe_acc; a.cells_cap = ctx->fe_cells_cap; a.tile_cnt = ctx->d_fe_tilecnt; a.tile_off = ctx->d_fe_tileoff; a.tiles_cap = ctx->fe_tiles_cap;
    a.X = b->X; a.x_off = reinterpret_cast<long long*>(b->x_offsets); a.x_cap = b->x_capacity; a.status = b->status;
    const long long np = (long long)b->rows * b->cols;
    const int gx = (int)std::min<long long>((np + 255) / 256, std::max(1, ctx->sm_count * 8 / std::max(1, std::min(b->n_frames, 8))));
    const dim3 gpix(gx, b->n_frames);
    const int gcell = ctx->sm_count * 8;
    fe_init_kernel<<<(b->n_frames * 6 + 255) / 256, 256, 0, stream>>>(a);
    fe_bbox_kernel<<<gpix, 256, 0, stream>>>(a);
    fe_layout_kernel<<<1, 32, 0, stream>>>(a);
    fe_clear_kernel<<<gcell, 256, 0, stream>>>(a);
    fe_accum_kernel<<<gpix, 256, 0, stream>>>(a);
    fe_count_kernel<<<gcell, 256, 0, stream>>>(a);
    fe_scan_kernel<<<1, 256, 0, stream>>>(a);
    fe_emit_kernel<<<gcell, 256, 0, stream>>>(a);
    CK(cudaGetLastError());
    ctx->info[5] = 8;
    return TDLO_OK;
}

extern "C" int tdlo_point_cloud_batched(tdlo_ctx* ctx, const tdlo_frontend_batch* b) {
    if (!ctx) return TDLO_ERR_INVALID;
    int rc = fe_check(ctx, b);
    if (rc) return rc;
    if (b->n_frames == 0) return TDLO_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t F = b->n_frames;
    const long long np = (long long)b->rows * b->cols;
    if ((long long)F * np > ctx->fe_pix_cap) {
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_fe_bgr) cudaFree(ctx->d_fe_bgr);
        if (ctx->d_fe_occl) cudaFree(ctx->d_fe_occl);
        if (ctx->d_fe_depth) cudaFree(ctx->d_fe_depth);
        ctx->d_fe_bgr = ctx->d_fe_occl = nullptr; ctx->d_fe_depth = nullptr; ctx->fe_pix_cap = 0;
        const size_t cap = (size_t)ctx->max_frames * np;
        CK(dalloc(&ctx->d_fe_bgr, cap * 3)); CK(dalloc(&ctx->d_fe_occl, cap * 3)); CK(dalloc(&ctx->d_fe_depth, cap));
        ctx->fe_pix_cap = (long long)cap;
    }
    if (!ctx->d_fe_proj) CK(dalloc(&ctx->d_fe_proj, (size_t)ctx->max_frames * 12));
    rc = fe_workspace(ctx);
    if (rc) return rc;
    H2D(ctx->d_fe_bgr, b->bgr, F * np * 3);
    H2D(ctx->d_fe_depth, b->depth, F * np * sizeof(unsigned short));
    if (b->occlusion_bgr) H2D(ctx->d_fe_occl, b->occlusion_bgr, F * np * 3);
    H2D(ctx->d_fe_proj, b->proj, F * 12 * sizeof(double));
    tdlo_frontend_batch d = *b;
    d.bgr = ctx->d_fe_bgr; d.depth = ctx->d_fe_depth; d.occlusion_bgr = b->occlusion_bgr ? ctx->d_fe_occl : nullptr; d.proj = ctx->d_fe_proj;
    d.X = ctx->d_X; d.x_offsets = reinterpret_cast<int64_t*>(ctx->d_xoff);
    d.x_capacity = std::min<long long>(b->x_capacity, ctx->max_points);
    d.status = ctx->d_fe_status;
    rc = tdlo_point_cloud_batched_device(ctx, &d, ctx->stream);
    if (rc) { cudaStreamSynchronize(ctx->stream); return rc; }
    D2H(b->x_offsets, ctx->d_xoff, (F + 1) * sizeof(long long));
    if (b->status) D2H(b->status, ctx->d_fe_status, F * sizeof(int));
    CK(cudaStreamSynchronize(ctx->stream));
    const long long total = b->x_offsets[F];
    if (total > 0) CK(cudaMemcpy(b->X, ctx->d_X, (size_t)total * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    return TDLO_OK;
}

// ---------------------------------------------------------------------------------------------
// evaluator frame error
// ---------------------------------------------------------------------------------------------
static int err_check(tdlo_ctx* ctx, const tdlo_err_batch* b) {
    if (!b) return fail(ctx, TDLO_ERR_INVALID, "null batch");
    if (b->n_frames < 0) return fail(ctx, TDLO_ERR_INVALID, "n_frames must be >= 0");
    if (b->n_track < 2 || b->n_track > TDLO_MAX_NODES || b->n_true < 2 || b->n_true > TDLO_MAX_NODES)
        return fail(ctx, TDLO_ERR_INVALID, "n_track / n_true must be in [2, %d]", TDLO_MAX_NODES);
    if (!b->Y_track || !b->Y_true || !b->error) return fail(ctx, TDLO_ERR_INVALID, "Y_track, Y_true, error are required");
    return TDLO_OK;
}

extern "C" int tdlo_tracking_error_batched_device(tdlo_ctx* ctx, const tdlo_err_batch* b, void* stream) {
    if (!ctx) return TDLO_ERR_INVALID;
    int rc = err_check(ctx, b);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    if (b->n_frames == 0) return TDLO_OK;
    tdlo_tracking_error_kernel<<<b->n_frames, 256, 0, (cudaStream_t)stream>>>(b->n_track, b->n_true, b->Y_track, b->Y_true, b->error);
    CK(cudaGetLastError());
    return TDLO_OK;
}

extern "C" int tdlo_tracking_error_batched(tdlo_ctx* ctx, const tdlo_err_batch* b) {
    if (!ctx) return TDLO_ERR_INVALID;
    int rc = err_check(ctx, b);
    if (rc) return rc;
    if (b->n_frames == 0) return TDLO_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t F = b->n_frames;
    double *dt = nullptr, *dr = nullptr, *de = nullptr;
    CK(dalloc(&dt, F * b->n_track * 3));
    cudaError_t e1 = dalloc(&dr, F * b->n_true * 3), e2 = dalloc(&de, F);
    if (e1 != cudaSuccess || e2 != cudaSuccess) { cudaFree(dt); cudaFree(dr); cudaFree(de); return fail(ctx, TDLO_ERR_NOMEM, "tracking error: out of device memory"); }
    cudaMemcpyAsync(dt, b->Y_track, F * b->n_track * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(dr, b->Y_true, F * b->n_true * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    tdlo_err_batch d = *b;
    d.Y_track = dt; d.Y_true = dr; d.error = de;
    rc = tdlo_tracking_error_batched_device(ctx, &d, ctx->stream);
    cudaMemcpyAsync(b->error, de, F * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t es = cudaStreamSynchronize(ctx->stream);
    cudaFree(dt); cudaFree(dr); cudaFree(de);
    if (rc) return rc;
    if (es != cudaSuccess) return fail(ctx, TDLO_ERR_CUDA, "tracking error: %s", cudaGetErrorString(es));
    return TDLO_OK;
}

// ---------------------------------------------------------------------------------------------
// sequence mode: visibility + tracking_step per frame, state carried on the device
// ---------------------------------------------------------------------------------------------
extern "C" int tdlo_track_sequences(tdlo_ctx* ctx, const tdlo_seq_batch* b, const tdlo_track_params* p) {
    if (!ctx) return TDLO_ERR_INVALID;
    if (!b || !p) return fail(ctx, TDLO_ERR_INVALID, "null batch/params");
    const int S = b->n_sequences, N = b->n_nodes, T = b->n_steps;
    if (S < 0 || S > ctx->max_frames) return fail(ctx, TDLO_ERR_INVALID, "n_sequences %d exceeds capacity %d", S, ctx->max_frames);
    if (N < 4 || N > ctx->max_nodes) return fail(ctx, TDLO_ERR_INVALID, "n_nodes %d outside [4,%d]", N, ctx->max_nodes);
    if (T < 0) return fail(ctx, TDLO_ERR_INVALID, "n_steps must be >= 0");
    if (!b->X || !b->x_offsets || !b->Y || !b->sigma2 || !b->geodesic_coord) return fail(ctx, TDLO_ERR_INVALID, "X, x_offsets, Y, sigma2, geodesic_coord are required");
    int rc = check_params(ctx, p->mu, p->beta, p->max_iter, p->prune_radius);
    if (rc) return rc;
    if (S == 0 || T == 0) return TDLO_OK;
    if (b->x_offsets[0] != 0) return fail(ctx, TDLO_ERR_INVALID, "x_offsets[0] must be 0");
    for (long long i = 0; i < (long long)T * S; i++) if (b->x_offsets[i + 1] < b->x_offsets[i]) return fail(ctx, TDLO_ERR_INVALID, "x_offsets not monotone");
    for (int t = 0; t < T; t++)
        if (b->x_offsets[(long long)(t + 1) * S] - b->x_offsets[(long long)t * S] > ctx->max_points)
            return fail(ctx, TDLO_ERR_INVALID, "step %d: %lld points exceed capacity %lld", t, (long long)(b->x_offsets[(long long)(t + 1) * S] - b->x_offsets[(long long)t * S]), ctx->max_points);
    CK(cudaSetDevice(ctx->device));
    if (!ctx->d_vdmin) {
        const size_t FM = ctx->max_frames, NM = ctx->max_nodes;
        CK(dalloc(&ctx->d_vdmin, FM * NM)); CK(dalloc(&ctx->d_vvis, FM * NM)); CK(dalloc(&ctx->d_vext, FM * NM));
    }
    const size_t SN = (size_t)S * N;
    if (b->proj) {
        if (b->rows < 1 || b->cols < 1 || b->rows > (1 << 20) || b->cols > (1 << 20)) return fail(ctx, TDLO_ERR_INVALID, "self-occlusion test: bad image size %d x %d", b->rows, b->cols);
        if (b->pixel_width < 2 || b->pixel_width > 2 * SO_MAX_RADIUS) return fail(ctx, TDLO_ERR_INVALID, "self-occlusion test: pixel_width %d outside [2, %d]", b->pixel_width, 2 * SO_MAX_RADIUS);
        rc = vis_workspace(ctx);
        if (rc) return rc;
        H2D(ctx->d_vproj, b->proj, (size_t)S * 12 * sizeof(double));
    }
    H2D(ctx->d_Y, b->Y, SN * 3 * sizeof(double));
    H2D(ctx->d_sigma2, b->sigma2, (size_t)S * sizeof(double));
    H2D(ctx->d_rest, b->geodesic_coord, SN * sizeof(double));
    std::vector<long long> xo((size_t)S + 1);
    for (int t = 0; t < T; t++) {
        const long long base = b->x_offsets[(long long)t * S];
        for (int s = 0; s <= S; s++) xo[s] = b->x_offsets[(long long)t * S + s] - base;
        // the offsets are consumed by kernels of THIS step only after the previous step's kernels (stream order); the
        // pageable-memory copy is staged before cudaMemcpyAsync returns, so `xo` can be reused in the next round
        H2D(ctx->d_xoff, xo.data(), ((size_t)S + 1) * sizeof(long long));
        if (xo[S] > 0) H2D(ctx->d_X, b->X + base * 3, (size_t)xo[S] * 3 * sizeof(double));
        tdlo_vis_batch v;
        memset(&v, 0, sizeof(v));
        v.n_frames = S; v.n_nodes = N;
        v.X = ctx->d_X; v.x_offsets = reinterpret_cast<const int64_t*>(ctx->d_xoff); v.Y = ctx->d_Y; v.node_coord = ctx->d_rest;
        v.visibility_threshold = p->visibility_threshold; v.d_vis = b->d_vis;
        v.visible = ctx->d_vvis; v.visible_offsets = reinterpret_cast<int64_t*>(ctx->d_visoff);
        v.visible_ext = ctx->d_vext; v.visible_ext_offsets = reinterpret_cast<int64_t*>(ctx->d_extoff);
        if (b->proj) { v.proj = ctx->d_vproj; v.rows = b->rows; v.cols = b->cols; v.pixel_width = b->pixel_width; }
        rc = vis_launch(ctx, &v, ctx->stream);
        if (rc) return rc;
        tdlo_track_batch d;
        memset(&d, 0, sizeof(d));
        d.n_frames = S; d.n_nodes = N;
        d.X = ctx->d_X; d.x_offsets = reinterpret_cast<const int64_t*>(ctx->d_xoff); d.Y = ctx->d_Y; d.sigma2 = ctx->d_sigma2;
        d.geodesic_coord = ctx->d_rest;
        d.visible = ctx->d_vvis; d.visible_offsets = reinterpret_cast<const int64_t*>(ctx->d_visoff);
        d.visible_ext = ctx->d_vext; d.visible_ext_offsets = reinterpret_cast<const int64_t*>(ctx->d_extoff);
        d.guide_nodes = ctx->d_guide; d.priors = ctx->d_priors_out; d.n_priors = ctx->d_npri_out;
        d.iters = ctx->d_iters; d.status = ctx->d_status; d.state = ctx->d_state;
        rc = tdlo_tracking_step_batched_device(ctx, &d, p, ctx->stream);
            if (rc) return rc;
        if (b->Y_traj) D2H(b->Y_traj + (size_t)t * SN * 3, ctx->d_Y, SN * 3 * sizeof(double));
        if (b->iters_traj) D2H(b->iters_traj + (size_t)t * S * 2, ctx->d_iters, (size_t)S * 2 * sizeof(int));
        if (b->status_traj) D2H(b->status_traj + (size_t)t * S, ctx->d_status, (size_t)S * sizeof(int));
    }
    D2H(b->Y, ctx->d_Y, SN * 3 * sizeof(double));
    D2H(b->sigma2, ctx->d_sigma2, (size_t)S * sizeof(double));
    return sync_and_check(ctx, ctx->stream);
}

extern "C" int tdlo_set_option(tdlo_ctx* ctx, int32_t option, double value) {
    if (!ctx) return TDLO_ERR_INVALID;
    switch (option) {
        case TDLO_OPT_CHUNK_POINTS: {
            const int c = (int)value;
            if (c != 0 && (c < 256 || c > (1 << 20) || (c % 32))) return fail(ctx, TDLO_ERR_INVALID, "chunk must be 0 (automatic) or a multiple of 32 in [256, 2^20]");
            ctx->tq_chunk = c; return TDLO_OK;
        }
        case TDLO_OPT_TRUNCATION:
            if (!(value >= 40.0 && value <= 745.2)) return fail(ctx, TDLO_ERR_INVALID, "truncation exponent must be in [40, 745.2]");
            ctx->tq_zcut = value; return TDLO_OK;
        case TDLO_OPT_TRUNCATION_REL:
            if (!(value >= 38.0 && value <= 745.2)) return fail(ctx, TDLO_ERR_INVALID, "relative truncation exponent must be in [38, 745.2]");
            ctx->tq_zrel = value; return TDLO_OK;
        case TDLO_OPT_INFLIGHT:
            if (value < 0) return fail(ctx, TDLO_ERR_INVALID, "inflight must be >= 0");
            ctx->tq_inflight = (int)value; return TDLO_OK;
        case TDLO_OPT_THREADS: {
            const int t = (int)value;
            if (t != 256) return fail(ctx, TDLO_ERR_INVALID, "threads must be 256 (the 224-thread x 3-CTA variant of round 1 no longer fits three CTAs per SM and was removed)");
            ctx->tq_threads = t; return TDLO_OK;
        }
        case TDLO_OPT_SOLVER:
            if (value != 0.0 && value != 1.0 && value != 2.0) return fail(ctx, TDLO_ERR_INVALID, "solver must be 0 (automatic), 1 (dense) or 2 (structured)");
            ctx->tq_solver = (int)value; return TDLO_OK;
        case TDLO_OPT_VOXEL_CELLS:
            if (value != 0.0 && (value < 4096.0 || value > (double)(1LL << 31))) return fail(ctx, TDLO_ERR_INVALID, "voxel cells must be 0 (automatic) or in [4096, 2^31]");
            ctx->fe_cells_opt = (long long)value; return TDLO_OK;
        case TDLO_OPT_WATCHDOG_MS:
            if (!(value >= 0.0)) return fail(ctx, TDLO_ERR_INVALID, "watchdog must be >= 0 ms (0 = off)");
            ctx->watchdog_ms = value; return TDLO_OK;
        default: return fail(ctx, TDLO_ERR_INVALID, "unknown option %d", option);
    }
}

// Development aid: enable (and read back / reset) the per-phase cycle counters of the kernel (slot meaning: header).
extern "C" int tdlo_profile_phases(tdlo_ctx* ctx, int32_t enable, uint64_t cycles[16]) {
    if (!ctx) return TDLO_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    if (cycles) {
        if (ctx->d_prof) CK(cudaMemcpy(cycles, ctx->d_prof_buf, 16 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        else memset(cycles, 0, 16 * sizeof(uint64_t));
    }
    CK(cudaMemset(ctx->d_prof_buf, 0, 16 * sizeof(uint64_t)));
    ctx->d_prof = enable ? ctx->d_prof_buf : nullptr;
    return TDLO_OK;
}
