// Eigen-facing drop-in for the reference tracker class.
//
// `class trackdlo` below has exactly the public interface of trackdlo/include/trackdlo.h:53-103
// (constructors :57-71, accessors :73-79, cpd_lle :81-95, tracking_step :97-102) so that
// trackdlo_node.cpp links unchanged (it default-constructs a global, copy-assigns a 12-argument
// instance, then calls initialize_nodes / initialize_geodesic_coord / tracking_step and the three
// getters: trackdlo_node.cpp:54,131,142-143,366-369).  Every numeric step runs on the GPU through
// the C ABI in trackdlo_b200.h with a batch of one frame; there is no CPU path in here.
//
// Build: compile the node with -I<repo>/include and link -ltrackdlo_b200 instead of compiling
// trackdlo/src/trackdlo.cpp.  Eigen is the only dependency of this header; a build without Eigen
// (this repo's CI) defines TRACKDLO_ADAPTER_MATRIX_HEADER to a minimal column-major MatrixXd.
//
// Where it goes: in trackdlo/include/trackdlo.h the class declaration (:53-130) is replaced by
// `#include <trackdlo_adapter.hpp>`, i.e. this header is included INSIDE the reference's `#ifndef TRACKDLO_H` guard
// (trackdlo.h:47-48, :131) -- it therefore carries a guard of its own and never tests TRACKDLO_H.
#ifndef TRACKDLO_B200_ADAPTER_HPP
#define TRACKDLO_B200_ADAPTER_HPP

#ifdef TRACKDLO_ADAPTER_MATRIX_HEADER
#include TRACKDLO_ADAPTER_MATRIX_HEADER
#else
#include <Eigen/Dense>
#endif

#include <cstdint>
#include <cstdio>
#include <memory>
#include <vector>

#include "trackdlo_b200.h"

#ifndef TRACKDLO_ADAPTER_LOG_INFO
#define TRACKDLO_ADAPTER_LOG_INFO(msg) ((void)0)            /* the reference uses ROS_INFO (trackdlo.cpp:931-981) */
#endif
#ifndef TRACKDLO_ADAPTER_LOG_ERROR
#define TRACKDLO_ADAPTER_LOG_ERROR(msg) std::fprintf(stderr, "[trackdlo_b200] %s\n", msg)
#endif

using Eigen::MatrixXd;

class trackdlo {
public:
    trackdlo() {}
    trackdlo(int num_of_nodes)
        : Y_(MatrixXd::Zero(num_of_nodes, 3)), guide_nodes_(MatrixXd::Zero(num_of_nodes, 3)), sigma2_(0.0), beta_(5.0),
          beta_pre_proc_(3.0), lambda_(1.0), lambda_pre_proc_(1.0), alpha_(0.0), k_vis_(0.0), mu_(0.05), max_iter_(50),
          tol_(0.00001), lle_weight_(1.0), visibility_threshold_(0.02) {}
    trackdlo(int num_of_nodes, double visibility_threshold, double beta, double lambda, double alpha, double k_vis,
             double mu, int max_iter, double tol, double beta_pre_proc, double lambda_pre_proc, double lle_weight)
        : Y_(MatrixXd::Zero(num_of_nodes, 3)), guide_nodes_(MatrixXd::Zero(num_of_nodes, 3)), sigma2_(0.0), beta_(beta),
          beta_pre_proc_(beta_pre_proc), lambda_(lambda), lambda_pre_proc_(lambda_pre_proc), alpha_(alpha), k_vis_(k_vis),
          mu_(mu), max_iter_(max_iter), tol_(tol), lle_weight_(lle_weight), visibility_threshold_(visibility_threshold) {}

    double get_sigma2() { return sigma2_; }
    MatrixXd get_tracking_result() { return Y_; }
    MatrixXd get_guide_nodes() { return guide_nodes_; }
    std::vector<MatrixXd> get_correspondence_pairs() { return correspondence_priors_; }
    void initialize_geodesic_coord(std::vector<double> geodesic_coord) {
        for (size_t i = 0; i < geodesic_coord.size(); i++) geodesic_coord_.push_back(geodesic_coord[i]);
    }
    void initialize_nodes(MatrixXd Y_init) { Y_ = Y_init; guide_nodes_ = Y_init; }
    void set_sigma2(double sigma2) { sigma2_ = sigma2; }

    // status mask (TDLO_ST_*) of the most recent cpd_lle / tracking_step; not part of the reference API
    int last_status() const { return last_status_; }

    bool cpd_lle(MatrixXd X_orig, MatrixXd& Y, double& sigma2, double beta, double lambda, double lle_weight, double mu,
                 int max_iter = 30, double tol = 0.0001, bool include_lle = true,
                 std::vector<MatrixXd> correspondence_priors = {}, double alpha = 0,
                 std::vector<int> visible_nodes = {}, double k_vis = 0, double visibility_threshold = 0.01) {
        const int Nn = (int)Y.rows();
        const int64_t Mp = (int64_t)X_orig.rows();
        if (!ensure_ctx(Nn, Mp)) return false;
        std::vector<double> X = to_rows(X_orig), Yr = to_rows(Y), W((size_t)Nn * 3);
        // the reference takes a prior list of ANY length (trackdlo.cpp:244-254): the whole list is passed on
        const int32_t n_pri = (int32_t)correspondence_priors.size();
        const int32_t pri_stride = n_pri > Nn ? n_pri : Nn;
        if (pri_stride > 4 * TDLO_MAX_NODES) { TRACKDLO_ADAPTER_LOG_ERROR("cpd_lle: correspondence prior list too long for the GPU path"); return false; }
        std::vector<double> pri((size_t)pri_stride * 4, 0.0);
        for (int32_t i = 0; i < n_pri; i++)
            for (int t = 0; t < 4; t++) pri[(size_t)i * 4 + t] = correspondence_priors[(size_t)i](0, t);
        const int64_t xoff[2] = {0, Mp};
        int32_t n_vis = (int32_t)visible_nodes.size(), iters = 0, status = 0;
        tdlo_cpd_batch b{};
        b.n_frames = 1; b.node_stride = Nn; b.X = X.data(); b.x_offsets = xoff; b.Y = Yr.data(); b.sigma2 = &sigma2;
        b.priors = pri.data(); b.n_priors = &n_pri; b.priors_stride = pri_stride; b.n_visible = &n_vis; b.W = W.data(); b.iters = &iters; b.status = &status;
        tdlo_cpd_params p{};
        p.beta = beta; p.lambda = lambda; p.lle_weight = lle_weight; p.mu = mu; p.tol = tol; p.alpha = alpha; p.k_vis = k_vis;
        p.visibility_threshold = visibility_threshold; p.prune_radius = 0.1; p.max_iter = max_iter; p.include_lle = include_lle ? 1 : 0;
        if (tdlo_cpd_lle_batched(ctx_->h, &b, &p) != TDLO_OK) { TRACKDLO_ADAPTER_LOG_ERROR(tdlo_last_error(ctx_->h)); return false; }
        last_status_ = status;
        from_rows(Yr, Y);
        report_status(status);
        if (status & TDLO_ST_NOT_CONVERGED) { TRACKDLO_ADAPTER_LOG_ERROR("optimization did not converge!"); return false; }  // trackdlo.cpp:434
        return true;
    }

    void tracking_step(MatrixXd X_orig, std::vector<int> visible_nodes, std::vector<int> visible_nodes_extended,
                       MatrixXd proj_matrix, int img_rows, int img_cols) {
        (void)proj_matrix; (void)img_rows; (void)img_cols;      // unused by the reference too (trackdlo.cpp:900-999)
        const int Nn = (int)Y_.rows();
        const int64_t Mp = (int64_t)X_orig.rows();
        correspondence_priors_.clear();
        if (!ensure_ctx(Nn, Mp)) return;
        std::vector<double> X = to_rows(X_orig), Yr = to_rows(Y_), guide((size_t)Nn * 3), pri((size_t)Nn * 8);
        std::vector<double> geo(geodesic_coord_.begin(), geodesic_coord_.end());
        geo.resize(Nn, geo.empty() ? 0.0 : geo.back());
        std::vector<int32_t> vis(visible_nodes.begin(), visible_nodes.end()), ext(visible_nodes_extended.begin(), visible_nodes_extended.end());
        const int64_t xoff[2] = {0, Mp}, voff[2] = {0, (int64_t)vis.size()}, eoff[2] = {0, (int64_t)ext.size()};
        int32_t n_pri = 0, iters[2] = {0, 0}, status = 0, state = -1;
        if (vis.empty()) vis.push_back(0);          // keep pointers valid; offsets still say "empty"
        if (ext.empty()) ext.push_back(0);
        tdlo_track_batch b{};
        b.n_frames = 1; b.n_nodes = Nn; b.X = X.data(); b.x_offsets = xoff; b.Y = Yr.data(); b.sigma2 = &sigma2_;
        b.geodesic_coord = geo.data(); b.visible = vis.data(); b.visible_offsets = voff; b.visible_ext = ext.data();
        b.visible_ext_offsets = eoff; b.guide_nodes = guide.data(); b.priors = pri.data(); b.n_priors = &n_pri;
        b.iters = iters; b.status = &status; b.state = &state;
        tdlo_track_params p{};
        p.visibility_threshold = visibility_threshold_; p.beta = beta_; p.lambda = lambda_; p.alpha = alpha_; p.k_vis = k_vis_;
        p.mu = mu_; p.tol = tol_; p.beta_pre_proc = beta_pre_proc_; p.lambda_pre_proc = lambda_pre_proc_;
        p.lle_weight = lle_weight_; p.prune_radius = 0.1; p.max_iter = max_iter_;
        if (tdlo_tracking_step_batched(ctx_->h, &b, &p) != TDLO_OK) { TRACKDLO_ADAPTER_LOG_ERROR(tdlo_last_error(ctx_->h)); return; }
        last_status_ = status;
        report_status(status);
        static const char* kState[] = {"All nodes visible / minor occlusion", "Mid-section occluded", "Tail occluded",
                                       "Head occluded", "Both ends occluded"};
        if (state >= 0 && state <= 4) { TRACKDLO_ADAPTER_LOG_INFO(kState[state]); }
        (void)kState;
        if (status & TDLO_ST_NOT_CONVERGED) TRACKDLO_ADAPTER_LOG_ERROR("optimization did not converge!");
        from_rows(Yr, Y_);
        const int V = (int)visible_nodes_extended.size();
        guide_nodes_ = MatrixXd::Zero(V, 3);
        for (int i = 0; i < V; i++) for (int d = 0; d < 3; d++) guide_nodes_(i, d) = guide[(size_t)i * 3 + d];
        for (int k = 0; k < n_pri; k++) {
            MatrixXd row = MatrixXd::Zero(1, 4);
            for (int t = 0; t < 4; t++) row(0, t) = pri[(size_t)k * 4 + t];
            correspondence_priors_.push_back(row);
        }
    }

private:
    struct Ctx {
        tdlo_ctx* h = nullptr;
        int nodes = 0;
        int64_t points = 0;
        ~Ctx() { if (h) tdlo_destroy(h); }
    };
    // copies of a tracker share one GPU context (the node copy-assigns the tracker once, trackdlo_node.cpp:131)
    std::shared_ptr<Ctx> ctx_;

    // inputs for which the reference itself has no defined result (it reads out of range / divides by zero): say so
    static void report_status(int status) {
        if (status & TDLO_ST_TOO_FEW_NODES) TRACKDLO_ADAPTER_LOG_ERROR("fewer than 4 (guide) nodes: registration skipped, nodes left unchanged");
        if (status & TDLO_ST_EMPTY_CLOUD) TRACKDLO_ADAPTER_LOG_ERROR("no point within 0.1 m of any node: registration skipped, nodes left unchanged");
        if (status & TDLO_ST_SINGULAR) TRACKDLO_ADAPTER_LOG_ERROR("singular / non-finite pivot in the M-step solve");
        if (status & TDLO_ST_TRAVERSE_UB) TRACKDLO_ADAPTER_LOG_ERROR("traverse_euclidean took a path on which the reference reads out of range; priors are defined but have no reference counterpart");
    }

    bool ensure_ctx(int nodes, int64_t points) {
        if (ctx_ && ctx_->h && ctx_->nodes >= nodes && ctx_->points >= points) return true;
        std::shared_ptr<Ctx> c = std::make_shared<Ctx>();
        c->nodes = nodes < 64 ? 64 : nodes;
        c->points = points < 65536 ? 65536 : 2 * points;
        if (tdlo_create(&c->h, 0, 1, c->nodes, c->points) != TDLO_OK) { TRACKDLO_ADAPTER_LOG_ERROR(tdlo_last_error(nullptr)); return false; }
        ctx_ = c;
        return true;
    }
    static std::vector<double> to_rows(const MatrixXd& m) {       // column-major MatrixXd -> row-major [n][cols]
        const size_t r = (size_t)m.rows(), c = (size_t)m.cols();
        std::vector<double> out(r * c);
        const double* d = m.data();
        for (size_t j = 0; j < c; j++) for (size_t i = 0; i < r; i++) out[i * c + j] = d[i + j * r];
        return out;
    }
    static void from_rows(const std::vector<double>& v, MatrixXd& m) {
        const size_t r = (size_t)m.rows(), c = (size_t)m.cols();
        double* d = m.data();
        for (size_t j = 0; j < c; j++) for (size_t i = 0; i < r; i++) d[i + j * r] = v[i * c + j];
    }

    MatrixXd Y_;
    MatrixXd guide_nodes_;
    double sigma2_ = 0.0;
    double beta_ = 5.0;
    double beta_pre_proc_ = 3.0;
    double lambda_ = 1.0;
    double lambda_pre_proc_ = 1.0;
    double alpha_ = 0.0;
    double k_vis_ = 0.0;
    double mu_ = 0.05;
    int max_iter_ = 50;
    double tol_ = 0.00001;
    double lle_weight_ = 1.0;
    std::vector<double> geodesic_coord_;
    std::vector<MatrixXd> correspondence_priors_;
    double visibility_threshold_ = 0.02;
    int last_status_ = 0;
};

#endif  // TRACKDLO_B200_ADAPTER_HPP
