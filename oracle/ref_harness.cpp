// TEST INFRASTRUCTURE ONLY -- C-ABI harness around the UNMODIFIED reference sources.
//
// oracle/Makefile (target `_ref`) compiles this file together with
//   /root/reference/trackdlo/src/trackdlo.cpp   and   /root/reference/trackdlo/src/utils.cpp
// from where they lie, against the reference's own headers (trackdlo/include/trackdlo.h, utils.h), into
// oracle/_ref/libtrackdlo_ref.so.  Nothing of the reference is copied into this repository.  The third-party
// headers those files name are absent from this image; they are satisfied by
//   oracle/ref_shim/eigen  -- an eager stand-in for the Eigen calls the two files make (see its header), or the REAL
//                             Eigen when one is available:  make -C oracle _ref EIGEN_INCLUDE=/usr/include/eigen3
//   oracle/ref_shim/stubs  -- empty ROS / OpenCV / PCL headers + the rosconsole macros routed to a log hook
// Every function below calls the reference's class exactly as trackdlo_node.cpp does (trackdlo_node.cpp:54,131,
// 142-143,366-369) and only converts row-major buffers <-> MatrixXd.  Used by tests/ (pins oracle/trackdlo_oracle.cpp
// and, on the GPU box, checks the CUDA path directly) and by bench.py's CPU legs; the product never loads it.

// standard headers first, so that the access-specifier trick below cannot touch them
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <thread>
#include <vector>
#include <signal.h>
#include <unistd.h>
#include <Eigen/Dense>
#include <Eigen/Geometry>

// traverse_euclidean / calc_LLE_weights are private members (trackdlo.h:123-128); the differential tests call them
// directly.  The sources stay untouched: only this translation unit sees the members as public (the class layout is
// unchanged, access specifiers do not affect it for this compiler).
#define private public
#include "trackdlo.h"
#undef private
#include "utils.h"

namespace {
thread_local std::vector<std::pair<int, std::string>> g_log;
}
namespace tdlo_ref_shim {
void log(int level, const std::string& msg) { g_log.emplace_back(level, msg); }
}

namespace {

using Eigen::MatrixXd;

MatrixXd from_rows(const double* p, int64_t n, int c) {
    MatrixXd m = MatrixXd::Zero(n, c);
    for (int64_t i = 0; i < n; i++) for (int d = 0; d < c; d++) m(i, d) = p[i * c + d];
    return m;
}
void to_rows(const MatrixXd& m, double* p) {
    for (int64_t i = 0; i < m.rows(); i++) for (int64_t d = 0; d < m.cols(); d++) p[i * m.cols() + d] = m(i, d);
}

// what one cpd_lle call logged: iterations executed (trackdlo.cpp:426 / :434); -1 if nothing was logged (max_iter==0)
int iters_from_log(size_t from, size_t to, int max_iter) {
    for (size_t i = from; i < to && i < g_log.size(); i++) {
        const std::string& s = g_log[i].second;
        const std::string key = "Iteration until convergence: ";
        if (s.compare(0, key.size(), key) == 0) return std::atoi(s.c_str() + key.size());
        if (s == "optimization did not converge!") return max_iter;
    }
    return max_iter <= 0 ? 0 : -1;
}

}  // namespace

extern "C" {

struct ref_cpd_params {        // same layout as oracle_cpd_params
    double beta, lambda, lle_weight, mu, tol, alpha, k_vis, visibility_threshold;
    int32_t max_iter, include_lle;
};
struct ref_track_params {      // same layout as oracle_track_params
    double visibility_threshold, beta, lambda, alpha, k_vis, mu, tol, beta_pre_proc, lambda_pre_proc, lle_weight;
    int32_t max_iter, pad_;
};

// 1 = the Eigen stand-in of oracle/ref_shim, 0 = real Eigen headers
int ref_uses_eigen_shim(void) {
#ifdef TDLO_REF_SHIM_EIGEN_DENSE
    return 1;
#else
    return 0;
#endif
}

// trackdlo::cpd_lle (trackdlo.h:81-95).  Returns converged (trackdlo.cpp:440).
int ref_cpd_lle(const double* X, int64_t n_points, double* Y, int32_t n_nodes, double* sigma2, const ref_cpd_params* p,
                const double* priors, int32_t n_priors, const int32_t* vis, int32_t n_vis, int32_t* iters_out) {
    g_log.clear();
    trackdlo t;
    MatrixXd Xm = from_rows(X, n_points, 3), Ym = from_rows(Y, n_nodes, 3);
    std::vector<MatrixXd> pr;
    for (int i = 0; i < n_priors; i++) pr.push_back(from_rows(priors + 4 * i, 1, 4));
    std::vector<int> v(vis, vis + (vis ? n_vis : 0));
    bool conv = t.cpd_lle(Xm, Ym, *sigma2, p->beta, p->lambda, p->lle_weight, p->mu, p->max_iter, p->tol, p->include_lle != 0,
                          pr, p->alpha, v, p->k_vis, p->visibility_threshold);
    to_rows(Ym, Y);
    if (iters_out) *iters_out = iters_from_log(0, g_log.size(), p->max_iter);
    return conv ? 1 : 0;
}

// The sequence trackdlo_node.cpp performs around one frame: construct (:131), initialize_nodes (:142),
// initialize_geodesic_coord (:143), [set_sigma2], tracking_step (:366), then the three getters (:367-369).
// state_out: 0 all visible / minor occlusion, 1 mid-section, 2 tail occluded, 3 head occluded, 4 both ends
// (from the ROS_INFO lines trackdlo.cpp:931-981).  Returns 0.
int ref_tracking_step(const double* X, int64_t n_points, double* Y, int32_t n_nodes, double* sigma2,
                      const double* geodesic_coord, const int32_t* vis, int32_t n_vis, const int32_t* vis_ext,
                      int32_t n_vis_ext, const ref_track_params* tp, double* guide_out, double* priors_out,
                      int32_t* n_priors_out, int32_t* iters_out /*[2]*/, int32_t* state_out) {
    g_log.clear();
    trackdlo tracker;
    tracker = trackdlo(n_nodes, tp->visibility_threshold, tp->beta, tp->lambda, tp->alpha, tp->k_vis, tp->mu, tp->max_iter,
                       tp->tol, tp->beta_pre_proc, tp->lambda_pre_proc, tp->lle_weight);
    tracker.initialize_nodes(from_rows(Y, n_nodes, 3));
    tracker.initialize_geodesic_coord(std::vector<double>(geodesic_coord, geodesic_coord + n_nodes));
    tracker.set_sigma2(*sigma2);
    std::vector<int> v(vis, vis + n_vis), ve(vis_ext, vis_ext + n_vis_ext);
    MatrixXd proj = MatrixXd::Zero(3, 4);
    tracker.tracking_step(from_rows(X, n_points, 3), v, ve, proj, 720, 1280);
    to_rows(tracker.get_tracking_result(), Y);
    *sigma2 = tracker.get_sigma2();
    if (guide_out) to_rows(tracker.get_guide_nodes(), guide_out);
    std::vector<MatrixXd> pr = tracker.get_correspondence_pairs();
    if (priors_out) for (size_t i = 0; i < pr.size(); i++) for (int k = 0; k < 4; k++) priors_out[i * 4 + k] = pr[i](0, k);
    if (n_priors_out) *n_priors_out = (int)pr.size();
    int state = -1;
    size_t state_pos = g_log.size();
    static const char* names[] = {"All nodes visible", "Minor occlusion", "Mid-section occluded", "Tail occluded",
                                  "Head occluded", "Both ends occluded"};
    static const int codes[] = {0, 0, 1, 2, 3, 4};
    for (size_t i = 0; i < g_log.size() && state < 0; i++)
        for (int k = 0; k < 6; k++) if (g_log[i].second == names[k]) { state = codes[k]; state_pos = i; break; }
    if (state_out) *state_out = state;
    if (iters_out) {
        iters_out[0] = iters_from_log(0, state_pos, tp->max_iter);
        iters_out[1] = iters_from_log(state_pos + 1, g_log.size(), tp->max_iter);
    }
    return 0;
}

// trackdlo::traverse_euclidean (trackdlo.cpp:584-898).  pairs_out [<= n_geo + 2][4].  Callers must not pass inputs for
// which the reference reads out of range (alignment 2 whose upward run reaches the end of the list, trackdlo.cpp:828).
int ref_traverse_euclidean(const double* geodesic_coord, int32_t n_geo, const double* guide, int32_t n_guide,
                           const int32_t* vis, int32_t n_vis, int32_t alignment, int32_t align_idx, double* pairs_out,
                           int32_t* n_pairs_out) {
    trackdlo t;
    std::vector<double> geo(geodesic_coord, geodesic_coord + n_geo);
    std::vector<int> v(vis, vis + n_vis);
    std::vector<MatrixXd> out = t.traverse_euclidean(geo, from_rows(guide, n_guide, 3), v, alignment, align_idx);
    for (size_t i = 0; i < out.size(); i++) for (int k = 0; k < 4; k++) pairs_out[i * 4 + k] = out[i](0, k);
    *n_pairs_out = (int)out.size();
    return 0;
}

// trackdlo::calc_LLE_weights(6, Y) (trackdlo.cpp:119-159) -> L [Nn][Nn] row-major, and H = (I-L)^T (I-L) (:237)
void ref_lle(const double* Y, int32_t n_nodes, double* L_out, double* H_out) {
    trackdlo t;
    MatrixXd L = t.calc_LLE_weights(6, from_rows(Y, n_nodes, 3));
    const int M = n_nodes;
    MatrixXd H = (MatrixXd::Identity(M, M) - L).transpose() * (MatrixXd::Identity(M, M) - L);
    if (L_out) to_rows(L, L_out);
    if (H_out) to_rows(H, H_out);
}

// line_sphere_intersection (utils.cpp:185-241): returns the number of intersections (0..2), points in out[2][3]
int ref_line_sphere_intersection(const double* A, const double* B, const double* centre, double radius, double* out) {
    std::vector<MatrixXd> r = line_sphere_intersection(from_rows(A, 1, 3), from_rows(B, 1, 3), from_rows(centre, 1, 3), radius);
    for (size_t i = 0; i < r.size(); i++) for (int k = 0; k < 3; k++) out[i * 3 + k] = r[i](0, k);
    return (int)r.size();
}

// The Eigen calls whose arithmetic is third-party (trackdlo.cpp:415 and :136-143), exposed so that tests can check the
// stand-in (or the real Eigen) against LAPACK: X = A.completeOrthogonalDecomposition().solve(B); inv = A.inverse();
// returns A.determinant().  Row-major n x n / n x k buffers.
void ref_eigen_cod_solve(const double* A, const double* B, int32_t n, int32_t k, double* X_out) {
    MatrixXd Am = from_rows(A, n, n), Bm = from_rows(B, n, k);
    MatrixXd W = Am.completeOrthogonalDecomposition().solve(Bm);
    to_rows(W, X_out);
}
double ref_eigen_inverse(const double* A, int32_t n, double* inv_out) {
    MatrixXd Am = from_rows(A, n, n);
    MatrixXd inv = Am.inverse();
    to_rows(inv, inv_out);
    return Am.determinant();
}

// pt2pt_dis / pt2pt_dis_sq (utils.cpp:13-19) on [n][3] inputs
double ref_pt2pt_dis(const double* a, const double* b, int32_t n) { return pt2pt_dis(from_rows(a, n, 3), from_rows(b, n, 3)); }
double ref_pt2pt_dis_sq(const double* a, const double* b, int32_t n) { return pt2pt_dis_sq(from_rows(a, n, 3), from_rows(b, n, 3)); }

}  // extern "C"
