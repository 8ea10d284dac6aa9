/* trackdlo_b200 -- C ABI of the B200-native TrackDLO registration path.
 *
 * Drop-in boundary for the reference's per-frame CPD/MCT EM registration:
 *   trackdlo::cpd_lle       (trackdlo/include/trackdlo.h:81-95,  trackdlo/src/trackdlo.cpp:161-441)
 *   trackdlo::tracking_step (trackdlo/include/trackdlo.h:97-102, trackdlo/src/trackdlo.cpp:900-999)
 * extended with a batch dimension over independent frames (BASELINE.json north_star).
 * The Eigen-facing `class trackdlo` with the reference's exact signature lives in
 * include/trackdlo_adapter.hpp and calls only the functions below.
 *
 * Conventions
 *   - plain C, no exceptions; every call returns TDLO_OK (0) or a negative error code and
 *     records a message retrievable with tdlo_last_error().
 *   - all matrices are row-major doubles: points [n][3] (xyz interleaved), nodes [n][3],
 *     priors [k][4] = {node_idx, x, y, z} exactly as trackdlo.cpp:247-250 reads them.
 *   - ragged per-frame arrays use CSR-style offsets ([n_frames+1], int64).
 *   - `*_device` entry points take DEVICE pointers and are stream-ordered (asynchronous);
 *     the plain entry points take HOST pointers, copy in, run, copy out and synchronise.
 *   - one context per host thread / GPU; contexts are not re-entrant (the reference class is
 *     single-threaded too, trackdlo_node.cpp:643).
 *   - there is no CPU fallback: every entry point fails with TDLO_ERR_CUDA when no sm_100 device
 *     is usable.
 */
#ifndef TRACKDLO_B200_H
#define TRACKDLO_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDLO_OK 0
#define TDLO_ERR_INVALID (-1)   /* bad argument / capacity exceeded */
#define TDLO_ERR_CUDA (-2)      /* CUDA runtime error (see tdlo_last_error) */
#define TDLO_ERR_NOMEM (-3)

/* per-frame status word written to `status[f]` (bit mask) */
#define TDLO_ST_NOT_CONVERGED 1  /* cpd_lle would return false (trackdlo.cpp:433-437)            */
#define TDLO_ST_SINGULAR 2       /* zero / non-finite pivot in the A W = B solve                */
#define TDLO_ST_TOO_FEW_NODES 4  /* fewer than 4 nodes: reference indexes out of range (:313-321) */
#define TDLO_ST_EMPTY_CLOUD 8    /* no point within prune_radius of any node (reference divides by 0) */
#define TDLO_ST_TRAVERSE_UB 16   /* traverse_euclidean / prior merge hit a path that is undefined
                                    behaviour in the reference; result is defined
                                    but has no reference counterpart                              */
#define TDLO_ST_PRE_NOT_CONVERGED 32 /* tracking_step: the pre-processing registration hit max_iter */
#define TDLO_ST_INVALID_INPUT 64 /* device-pointer entry points: the frame's offsets / node count / visibility lists would
                                    overrun the context's capacity or index outside [0, n_nodes): frame not run, outputs
                                    untouched (the host-pointer entry points reject such a batch with TDLO_ERR_INVALID) */

#define TDLO_MAX_NODES 256

typedef struct tdlo_ctx tdlo_ctx;

/* Arguments of trackdlo::cpd_lle (trackdlo.h:81-95) that are not per-frame data. */
typedef struct tdlo_cpd_params {
    double beta;                 /* MCT kernel width                                   */
    double lambda;               /* MCT weight                                         */
    double lle_weight;           /* gamma, used iff include_lle                        */
    double mu;                   /* outlier ratio, 0 < mu < 1                          */
    double tol;                  /* mean node displacement threshold (trackdlo.cpp:424) */
    double alpha;                /* correspondence-prior weight                        */
    double k_vis;                /* visibility strength (trackdlo.cpp:358-379)          */
    double visibility_threshold; /* tau_vis (trackdlo.cpp:291)                          */
    double prune_radius;         /* 0.1 in the reference (trackdlo.cpp:190)             */
    int32_t max_iter;
    int32_t include_lle;
} tdlo_cpd_params;

/* Constructor arguments of class trackdlo (trackdlo.h:59-71). */
typedef struct tdlo_track_params {
    double visibility_threshold, beta, lambda, alpha, k_vis, mu, tol;
    double beta_pre_proc, lambda_pre_proc, lle_weight;
    double prune_radius;         /* 0.1 */
    int32_t max_iter;
    int32_t reserved;
} tdlo_track_params;

/* A batch of independent cpd_lle problems.  Pointers are all-host or all-device depending on
 * the entry point.  Optional arrays may be NULL. */
typedef struct tdlo_cpd_batch {
    int32_t n_frames;
    int32_t node_stride;        /* row stride (in nodes) of Y / priors / W / H; >= every n_nodes[f] */
    const double* X;            /* [x_offsets[n_frames]][3]  X_orig of every frame, concatenated   */
    const int64_t* x_offsets;   /* [n_frames+1]                                                     */
    const int32_t* n_nodes;     /* [n_frames] or NULL (= node_stride for all frames)                */
    double* Y;                  /* [n_frames][node_stride][3]  in: Y, out: registered Y             */
    double* sigma2;             /* [n_frames] in/out; 0 selects the data-driven init (:271-273)     */
    const double* priors;       /* [n_frames][priors_stride][4] or NULL                             */
    const int32_t* n_priors;    /* [n_frames] or NULL; 0 <= n_priors[f] <= priors_stride            */
    const int32_t* n_visible;   /* [n_frames] or NULL: visible_nodes.size() (cpd_lle only uses the
                                   size of that vector, trackdlo.cpp:358)                           */
    const double* H;            /* optional [n_frames][node_stride][node_stride] LLE matrix
                                   H=(I-L)^T(I-L); NULL -> computed on the device (:236-237)        */
    double* W;                  /* optional out [n_frames][node_stride][3] last W (:415)           */
    int32_t* iters;             /* optional out [n_frames] EM iterations executed                   */
    int32_t* status;            /* optional out [n_frames] TDLO_ST_* mask                           */
    int32_t priors_stride;      /* rows per frame in `priors`; 0 = node_stride.  The reference accepts a prior list of
                                   any length (later rows overwrite earlier ones with the same node index,
                                   trackdlo.cpp:244-254): pass a larger stride for lists longer than the node count  */
    int32_t reserved;
} tdlo_cpd_batch;

/* A batch of independent tracking_step problems (one tracker object each). */
typedef struct tdlo_track_batch {
    int32_t n_frames;
    int32_t n_nodes;             /* nodes per tracker (Y_.rows()); same for the whole batch         */
    const double* X;             /* [x_offsets[n_frames]][3]                                         */
    const int64_t* x_offsets;    /* [n_frames+1]                                                     */
    double* Y;                   /* [n_frames][n_nodes][3]  in: Y_, out: tracking result             */
    double* sigma2;              /* [n_frames] in/out sigma2_                                        */
    const double* geodesic_coord;/* [n_frames][n_nodes] rest arc-lengths (initialize_geodesic_coord) */
    const int32_t* visible;      /* ragged visible_nodes                                             */
    const int64_t* visible_offsets;      /* [n_frames+1]                                             */
    const int32_t* visible_ext;  /* ragged visible_nodes_extended (sorted ascending)                 */
    const int64_t* visible_ext_offsets;  /* [n_frames+1]                                             */
    const double* H_pre;         /* optional [n_frames][n_nodes][n_nodes] LLE matrix of the guide
                                    nodes for the pre-processing call (leading |vis_ext|^2 block,
                                    row stride n_nodes); NULL -> computed on the device              */
    double* guide_nodes;         /* optional out [n_frames][n_nodes][3] (first |vis_ext| rows valid) */
    double* priors;              /* optional out [n_frames][2*n_nodes][4] correspondence_priors_     */
    int32_t* n_priors;           /* optional out [n_frames]                                          */
    int32_t* iters;              /* optional out [n_frames][2] {pre-proc, main} EM iterations        */
    int32_t* status;             /* optional out [n_frames]                                          */
    int32_t* state;              /* optional out [n_frames] 0 all visible,1 mid occluded,2 tail occluded,
                                    3 head occluded,4 both ends occluded (trackdlo.cpp:929-995)      */
    double* packed_results;      /* optional out [n_frames][3*n_nodes + 4]: per frame one contiguous record
                                    {Y[n_nodes][3], sigma2, iters_pre, iters_main, status} written by the kernel's
                                    epilogue -- the payload a multi-GPU caller all-gathers in ONE collective     */
} tdlo_track_batch;

/* Creates a context on CUDA device `device`.  Capacities bound later batches:
 * max_frames frames per call, max_nodes <= TDLO_MAX_NODES nodes per frame,
 * max_points_total points summed over a batch. */
int tdlo_create(tdlo_ctx** out, int device, int32_t max_frames, int32_t max_nodes, int64_t max_points_total);
void tdlo_destroy(tdlo_ctx* ctx);
const char* tdlo_last_error(const tdlo_ctx* ctx);   /* ctx may be NULL: last create() error */
const char* tdlo_version(void);

/* trackdlo::cpd_lle over a batch.  Host pointers; synchronous. */
int tdlo_cpd_lle_batched(tdlo_ctx* ctx, const tdlo_cpd_batch* batch, const tdlo_cpd_params* params);
/* Same with device pointers; asynchronous on `stream` (a cudaStream_t, NULL = default stream).
 * PRECONDITIONS the host cannot check on device memory: x_offsets ascending from >= 0 with x_offsets[n_frames] <=
 * max_points_total of the context, n_nodes[f] <= node_stride, n_priors[f] <= priors_stride.  The kernel re-checks them
 * per frame: a violating frame is not run and gets TDLO_ST_INVALID_INPUT (n_priors is clamped). */
int tdlo_cpd_lle_batched_device(tdlo_ctx* ctx, const tdlo_cpd_batch* batch, const tdlo_cpd_params* params,
                                void* stream);

/* trackdlo::tracking_step over a batch.  Host pointers; synchronous. */
int tdlo_tracking_step_batched(tdlo_ctx* ctx, const tdlo_track_batch* batch, const tdlo_track_params* params);
/* Same with device pointers; asynchronous on `stream`.  Same per-frame re-check as above, plus: both visibility
 * lists at most n_nodes long, entries in [0, n_nodes), visible_ext strictly ascending. */
int tdlo_tracking_step_batched_device(tdlo_ctx* ctx, const tdlo_track_batch* batch,
                                      const tdlo_track_params* params, void* stream);

/* Visibility front-end that feeds tracking_step (trackdlo/src/trackdlo_node.cpp:254-277, 346-360; the self-occlusion
 * raster :280-343 is not included: every node counts as not self-occluded).  For every frame: shortest distance of
 * each node of Y to the frame's points (100000 if there is none, as in the reference), visible_nodes = nodes with
 * distance <= visibility_threshold (ascending), visible_nodes_extended = gaps filled where the rest arc length between
 * consecutive visible nodes is <= d_vis.  The CSR outputs are exactly the visibility inputs of tdlo_track_batch. */
typedef struct tdlo_vis_batch {
    int32_t n_frames;
    int32_t n_nodes;
    const double* X;              /* [x_offsets[n_frames]][3]                                  */
    const int64_t* x_offsets;     /* [n_frames+1]                                              */
    const double* Y;              /* [n_frames][n_nodes][3] current node estimate (Y^{t-1})     */
    const double* node_coord;     /* [n_frames][n_nodes] converted_node_coord (rest arc lengths) */
    double visibility_threshold;  /* trackdlo_node.cpp:319                                      */
    double d_vis;                 /* trackdlo_node.cpp:354                                      */
    double* dmin;                 /* optional out [n_frames][n_nodes]                           */
    int32_t* visible;             /* out, capacity n_frames*n_nodes                             */
    int64_t* visible_offsets;     /* out [n_frames+1]                                           */
    int32_t* visible_ext;         /* out, capacity n_frames*n_nodes                             */
    int64_t* visible_ext_offsets; /* out [n_frames+1]                                           */
    /* Optional self-occlusion test (trackdlo_node.cpp:280-343): with proj != NULL a node is only visible if, in addition,
     * its pixel is not covered by the thick lines (cv::line, thickness dlo_pixel_width) of the edges nearer to the camera that
     * the reference has drawn before it visits the node.  Evaluated without a raster, bit-exact with OpenCV 4.x's cv::line
     * (oracle/raster.py, pinned against cv2).  A zero-initialised tail keeps the old behaviour (every node not self-occluded). */
    const double* proj;           /* [n_frames][12] row-major 3x4 projection matrices (camera_info P), or NULL = off */
    int32_t rows, cols;           /* image size of the camera (mask.rows, mask.cols)                                  */
    int32_t pixel_width;          /* dlo_pixel_width (launch default 40); 2..512                                     */
    int32_t reserved;
    int32_t* not_self_occluded;   /* optional out [n_frames][n_nodes]: 1 = the reference's not_self_occluded_nodes      */
} tdlo_vis_batch;
/* Host pointers; synchronous. */
int tdlo_visibility_batched(tdlo_ctx* ctx, const tdlo_vis_batch* batch);
/* Device pointers; asynchronous on `stream` (no host read-back: the slice table is built on the device) -- chain it in
 * front of tdlo_tracking_step_batched_device. */
int tdlo_visibility_batched_device(tdlo_ctx* ctx, const tdlo_vis_batch* batch, void* stream);

/* Perception front-end that produces the tracker's input cloud (trackdlo/src/trackdlo_node.cpp:159-242), for a
 * batch of independent frames: BGR image -> HSV -> colour band(s) (cv::inRange, or the node's four-band color_thresholding)
 * -> AND with the grey occlusion image -> masked pixels + depth (uint16, millimetres) -> points through the projection
 * matrix (float32, like pcl::PointXYZRGB) -> pcl::VoxelGrid centroid down-sampling -> double.  Output: the points of all
 * frames concatenated (ascending voxel index inside a frame, PCL's order) + CSR offsets = the X / x_offsets inputs of
 * tdlo_visibility_batched(_device) and tdlo_tracking_step_batched(_device).  The OpenCV stages are bit-exact with OpenCV
 * (pinned against cv2 over all 2^24 colours); voxel membership and order follow PCL 1.10, the centroid is the exact mean
 * rounded once to float32 (PCL sums in float32 in an unspecified order): see oracle/frontend.py. */
#define TDLO_FE_EMPTY 1      /* status: no pixel of the frame passed the mask                                     */
#define TDLO_FE_GRID 2       /* status: the voxel grid over the masked points exceeds INT_MAX cells (PCL's own bail-out)
                                or the context's grid workspace (TDLO_OPT_VOXEL_CELLS): frame not down-sampled, 0 points */
#define TDLO_FE_CAPACITY 4   /* status: X is full: this frame and all later ones got 0 points                      */
typedef struct tdlo_frontend_batch {
    int32_t n_frames;
    int32_t rows, cols;              /* image size, the same for the whole batch                                  */
    int32_t multi_color;             /* 0: one band [hsv_lower, hsv_upper]; 1: color_thresholding (trackdlo_node.cpp:88-119) */
    const uint8_t* bgr;              /* [n_frames][rows][cols][3]                                                  */
    const uint16_t* depth;           /* [n_frames][rows][cols] millimetres (trackdlo_node.cpp:216)                  */
    const uint8_t* occlusion_bgr;    /* optional [n_frames][rows][cols][3] (the /mask_with_occlusion image, :172-180) */
    const double* proj;              /* [n_frames][12] row-major 3x4 projection matrix (fx, fy, cx, cy are read)    */
    int32_t hsv_lower[3], hsv_upper[3];   /* launch/trackdlo.launch:8-10 defaults: 90 90 30 / 130 255 255           */
    double leaf_size;                /* downsample_leaf_size, 0.008                                                */
    double* X;                       /* out [x_capacity][3]                                                         */
    int64_t* x_offsets;              /* out [n_frames+1]                                                            */
    int64_t x_capacity;              /* points X can hold                                                          */
    int32_t* status;                 /* optional out [n_frames] TDLO_FE_* mask                                      */
} tdlo_frontend_batch;
int tdlo_point_cloud_batched(tdlo_ctx* ctx, const tdlo_frontend_batch* batch);                        /* host pointers, synchronous */
int tdlo_point_cloud_batched_device(tdlo_ctx* ctx, const tdlo_frontend_batch* batch, void* stream);   /* device pointers, stream-ordered */

/* Evaluator frame error (trackdlo/src/evaluator.cpp:233-283, 333-341): for every frame the mean distance of
 * the nodes of Y_track to the nearest segment of the polyline Y_true, symmetrised ((E1 + E2) / 2). */
typedef struct tdlo_err_batch {
    int32_t n_frames;
    int32_t n_track;          /* nodes of every tracked polyline (<= TDLO_MAX_NODES, >= 2) */
    int32_t n_true;           /* nodes of every ground-truth polyline (<= TDLO_MAX_NODES, >= 2) */
    int32_t reserved;
    const double* Y_track;    /* [n_frames][n_track][3] */
    const double* Y_true;     /* [n_frames][n_true][3]  */
    double* error;            /* out [n_frames]         */
} tdlo_err_batch;
int tdlo_tracking_error_batched(tdlo_ctx* ctx, const tdlo_err_batch* batch);                      /* host pointers, synchronous */
int tdlo_tracking_error_batched_device(tdlo_ctx* ctx, const tdlo_err_batch* batch, void* stream);  /* device pointers */

/* Sequence mode: S independent trackers advanced over T consecutive frames WITHOUT returning to the host
 * between frames.  Per step t and sequence s, exactly what trackdlo_node.cpp does per callback: visibility lists from
 * Y^{t-1} and the step's cloud (tdlo_visibility_batched semantics, visibility_threshold = params->visibility_threshold),
 * then tracking_step, which leaves Y^{t} and sigma2 in place for step t+1 (trackdlo.cpp:998).  Host pointers; every copy
 * and kernel of all T steps is queued on one stream without a host synchronisation in between (the host runs ahead of the
 * device), one synchronisation at the end. */
typedef struct tdlo_seq_batch {
    int32_t n_sequences;
    int32_t n_nodes;
    int32_t n_steps;
    const double* X;              /* clouds of all steps, step-major then sequence: [x_offsets[T*S]][3]            */
    const int64_t* x_offsets;     /* [n_steps * n_sequences + 1]                                                   */
    double* Y;                    /* [S][n_nodes][3] in: initial nodes (initialize_nodes), out: nodes after step T  */
    double* sigma2;               /* [S] in/out                                                                    */
    const double* geodesic_coord; /* [S][n_nodes] rest arc lengths (initialize_geodesic_coord / converted_node_coord) */
    double d_vis;                 /* trackdlo_node.cpp:354                                                         */
    double* Y_traj;               /* optional out [T][S][n_nodes][3] tracking result after every step              */
    int32_t* iters_traj;          /* optional out [T][S][2]                                                        */
    int32_t* status_traj;         /* optional out [T][S]                                                           */
    /* optional self-occlusion test of the visibility lists (tdlo_vis_batch): a zero-initialised tail = off */
    const double* proj;           /* [S][12] row-major 3x4 projection matrix of every sequence's camera, or NULL      */
    int32_t rows, cols;           /* image size                                                                    */
    int32_t pixel_width;          /* dlo_pixel_width, 2..512                                                       */
    int32_t reserved;
} tdlo_seq_batch;
int tdlo_track_sequences(tdlo_ctx* ctx, const tdlo_seq_batch* batch, const tdlo_track_params* params);

/* Multi-rank helper for hosts without Python (frames shard across GPUs with no data-path collective; the only exchange is the
 * gather of the tracked nodes at the end).  All-gathers the packed result records of a sharded batch over an NCCL communicator
 * the CALLER created: `nccl_comm` is its ncclComm_t (passed as void* so that this header needs no nccl.h).  `packed_all` is
 * device memory [world * frames_per_rank][3 n_nodes + 4]; this rank's records must already sit at
 * packed_all + rank * frames_per_rank * (3 n_nodes + 4) -- point tdlo_track_batch::packed_results there and the persistent
 * kernel writes them in place.  Queued on `stream` behind the tracking call (one ncclAllGather, in place).  libnccl.so.2 is
 * loaded on first use (dlopen): this library does not link against NCCL.  Returns TDLO_ERR_CUDA with tdlo_last_error set if
 * NCCL is not available or reports an error. */
int tdlo_all_gather_packed(tdlo_ctx* ctx, void* nccl_comm, double* packed_all, int32_t frames_per_rank, int32_t n_nodes,
                           int32_t rank, void* stream);

/* Waits for the most recent *_device call of this context and reports what only the device knows: TDLO_ERR_CUDA if the
 * persistent kernel's watchdog gave up (a CTA waited longer than TDLO_OPT_WATCHDOG_MS for a task; results invalid),
 * TDLO_OK otherwise.  The host-pointer entry points do this themselves. */
int tdlo_synchronize(tdlo_ctx* ctx);

/* Launch geometry of the most recent call (for benchmarks / profiling):
 * info[0]=cluster size, [1]=CTAs launched, [2]=threads per CTA, [3]=dynamic smem bytes,
 * [4]=points per tile, [5]=kernels launched by that call, [6]=resident CTAs per SM, [7]=SM count. */
int tdlo_last_launch_info(const tdlo_ctx* ctx, int32_t info[8]);

/* Development aid: enables/disables the kernel's per-phase cycle counters and returns + resets the totals accumulated
 * since the previous call.  Task-queue engine: cycles[0..9] = {queue wait, prune, visibility pre-pass, E-step, start_call,
 * wave glue, M-step gather+assemble, solve, update, finish_call}, [10..15] = {E-step tasks, tiles, sum of window widths,
 * row blocks, per-warp tile-loop cycles, end-of-task reduction cycles}.  Synchronises. */
int tdlo_profile_phases(tdlo_ctx* ctx, int32_t enable, uint64_t cycles[16]);

/* Engine options (the engine: ONE persistent launch per call; the E-step of every frame is cut into chunk tasks on a
 * global ticket queue that any SM may run; the CTA finishing a frame's last chunk runs its M-step).
 *  TDLO_OPT_CHUNK_POINTS  raw points per chunk task; 0 (default) = automatic from the context's point
 *                         capacity: 256 / 512 for small contexts (one live sequence), 1024, or 2048 / 4096 for large batches (results are
 *                         bit-deterministic for a given chunk size).
 *  TDLO_OPT_TRUNCATION    z_cut: affinity entries exp(-z) with z > z_cut are skipped.  745.2 skips only entries
 *                         that are exactly 0 in the reference (double underflow); the default 100 skips entries
 *                         below 3.8e-44 of the column maximum, i.e. far below one ulp of every sum they enter.
 *  TDLO_OPT_TRUNCATION_REL  z_rel: additionally, entries more than exp(-z_rel) below the LARGEST entry of their column (the point's
 *                         nearest node) are skipped.  Default 45: what is dropped is below 3e-20 of the column sum it would
 *                         enter (half an ulp is 1.1e-16), so every sum is unchanged to the last bit or two; 745.2 = off.
 *  TDLO_OPT_INFLIGHT      frames in flight at once (0 = automatic).
 *  TDLO_OPT_THREADS       threads per CTA: 256 (2 CTAs/SM, 128 registers) is the only variant left; kept for ABI compatibility.
 *  TDLO_OPT_SOLVER        the M-step solve (trackdlo.cpp:394-417).  G is the Matern-3/2 covariance of the nodes' arc
 *                         lengths, so (S G + lambda sigma2 I) W = B and T = Y0 + G W can be computed in O(Nn) without forming
 *                         G or A: with S = diag(P1 + alpha J) by a state-space (Kalman filter + adjoint) recursion over the
 *                         nodes; with the LLE term S = diag(..) + sigma2 gamma H by a banded LDL^T of the information form
 *                         (c K + P^T S P) z = P^T B on the joint values (f, f') of the process, K its block-tridiagonal
 *                         precision.  Both are MORE accurate than a dense solve of the ill-conditioned A (1e-13 vs 1e-9
 *                         against a 50-digit solve: profiles/r2_kalman_solver_accuracy.txt, r2_banded_solver_accuracy.txt).
 *                         0 (default) and 2 = structured; 1 = dense always (register-resident Gauss-Jordan for Nn <= 64,
 *                         blocked Cholesky with FP64 tensor-core MMAs / pivoted elimination above).  A caller-supplied H and
 *                         a negative alpha always solve densely.
 *  TDLO_OPT_VOXEL_CELLS   cells of the front-end's voxel-grid workspace, summed over a batch (default: 2^19 per frame of the
 *                         context, at least 2^22, at most 2^25; 32 B each).  Takes effect at the next front-end call.
 *  TDLO_OPT_WATCHDOG_MS   a CTA of the persistent kernel that waits longer than this for its next task (or for a frame's
 *                         upload) abandons the launch and the call returns TDLO_ERR_CUDA instead of hanging the caller
 *                         (default 20000; 0 = never). */
#define TDLO_OPT_CHUNK_POINTS 2
#define TDLO_OPT_TRUNCATION 3
#define TDLO_OPT_INFLIGHT 4
#define TDLO_OPT_THREADS 5
#define TDLO_OPT_WATCHDOG_MS 6
#define TDLO_OPT_SOLVER 7
#define TDLO_OPT_VOXEL_CELLS 8
#define TDLO_OPT_TRUNCATION_REL 9
int tdlo_set_option(tdlo_ctx* ctx, int32_t option, double value);

#ifdef __cplusplus
}
#endif
#endif /* TRACKDLO_B200_H */
