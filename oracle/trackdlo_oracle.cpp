// TEST INFRASTRUCTURE ONLY -- CPU oracle for the TrackDLO CPD/MCT EM registration path.
//
// This file is a dependency-free C++17 *restatement* of the reference algorithm
//   trackdlo/src/trackdlo.cpp:92-159   (LLE weights)
//   trackdlo/src/trackdlo.cpp:161-441  (cpd_lle)
//   trackdlo/src/trackdlo.cpp:584-898  (traverse_euclidean)
//   trackdlo/src/trackdlo.cpp:900-999  (tracking_step)
//   trackdlo/src/utils.cpp:13-19,172-241 (pt2pt_dis, isBetween, line_sphere_intersection)
// written from SURVEY.md Appendix A/B.  It is used by tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py and by nothing else: the product
// (trackdlo_b200/) never links, imports or calls it.
//
// PINNING: the reference ships no golden vectors / known-answer tests.  This restatement is pinned against the
// reference's own statements instead: oracle/Makefile target `_ref` compiles the UNMODIFIED trackdlo.cpp + utils.cpp
// behind ref_harness.cpp, and tests/test_ref_pin.py compares the two on every golden input, random option sweeps, all
// five tracking_step states and all traversal alignments (agreement 1e-12..1e-14; integer outputs identical).
// What that build cannot pin is the third-party arithmetic outside /root/reference -- Eigen 3.3.7 (docs/RUN.md:10),
// absent from this image -- which both this file and the _ref build's stand-in (ref_shim/eigen) restate:
//   * completeOrthogonalDecomposition().solve (trackdlo.cpp:415)  -> column-pivoted Householder QR solve
//     (same algorithm class; identical for full-rank A; checked against LAPACK in test_ref_pin.py)
//   * .inverse()/.determinant() on the LLE Gram matrices (trackdlo.cpp:136-143) -> partial-pivot LU.  Those Gram
//     matrices are rank-3 6x6, so their inverse is rounding noise; no restatement can reproduce
//     Eigen's bits there.  With a real Eigen: make -C oracle -B _ref EIGEN_INCLUDE=/usr/include/eigen3.
// The oracle is cross-checked against an independent NumPy/LAPACK twin (oracle/numpy_twin.py).
//
// Matrices at the C boundary are row-major [rows][3] doubles (NumPy default).
//
// Deliberate deviations from undefined behaviour in the reference are marked "UB:" below.

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <array>
#include <algorithm>

namespace {

// column-major dense matrix (same element order as Eigen::MatrixXd)
struct Mat {
    int r = 0, c = 0;
    std::vector<double> d;
    Mat() {}
    Mat(int r_, int c_, double v = 0.0) : r(r_), c(c_), d((size_t)r_ * c_, v) {}
    double& operator()(int i, int j) { return d[(size_t)i + (size_t)j * r]; }
    double operator()(int i, int j) const { return d[(size_t)i + (size_t)j * r]; }
};

struct Vec3 { double x, y, z; };

inline Vec3 row3(const Mat& m, int i) { return {m(i, 0), m(i, 1), m(i, 2)}; }
inline double sqdist(Vec3 a, Vec3 b) {
    double dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return dx * dx + dy * dy + dz * dz;
}
inline double dist(Vec3 a, Vec3 b) { return std::sqrt(sqdist(a, b)); }

// ---------------------------------------------------------------------------------------------
// Dense solve: column-pivoted Householder QR (stand-in for Eigen COD, trackdlo.cpp:415).
// Solves A W = B for square A (n x n), B (n x k).  Rank is truncated with Eigen's default
// threshold (eps * n * max|R_ii|); for a rank-deficient A this returns the basic solution, not
// COD's minimum-norm one (never reached on this path: A = diag(P1) G + lambda*sigma2*I, c > 0).
// ---------------------------------------------------------------------------------------------
void colpiv_qr_solve(Mat A, Mat B, Mat& W) {
    const int n = A.r, k = B.c;
    std::vector<int> perm(n);
    std::vector<double> cn(n);
    for (int j = 0; j < n; j++) {
        perm[j] = j;
        double s = 0;
        for (int i = 0; i < n; i++) s += A(i, j) * A(i, j);
        cn[j] = s;
    }
    std::vector<double> v(n);
    double maxdiag = 0;
    int rank = n;
    for (int j = 0; j < n; j++) {
        // recompute remaining squared column norms (no downdating -> no cancellation issues)
        int best = j;
        double bestv = -1;
        for (int jj = j; jj < n; jj++) {
            double s = 0;
            for (int i = j; i < n; i++) s += A(i, jj) * A(i, jj);
            cn[jj] = s;
            if (s > bestv) { bestv = s; best = jj; }
        }
        if (best != j) {
            for (int i = 0; i < n; i++) std::swap(A(i, j), A(i, best));
            std::swap(perm[j], perm[best]);
            std::swap(cn[j], cn[best]);
        }
        double normx = std::sqrt(cn[j]);
        if (normx == 0.0) { rank = j; break; }
        double alpha = A(j, j) > 0 ? -normx : normx;
        // v = x - alpha e1
        for (int i = j; i < n; i++) v[i] = A(i, j);
        v[j] -= alpha;
        double vtv = 0;
        for (int i = j; i < n; i++) vtv += v[i] * v[i];
        if (vtv > 0) {
            for (int jj = j + 1; jj < n; jj++) {
                double s = 0;
                for (int i = j; i < n; i++) s += v[i] * A(i, jj);
                s = 2 * s / vtv;
                for (int i = j; i < n; i++) A(i, jj) -= s * v[i];
            }
            for (int jj = 0; jj < k; jj++) {
                double s = 0;
                for (int i = j; i < n; i++) s += v[i] * B(i, jj);
                s = 2 * s / vtv;
                for (int i = j; i < n; i++) B(i, jj) -= s * v[i];
            }
        }
        A(j, j) = alpha;
        for (int i = j + 1; i < n; i++) A(i, j) = 0;
        maxdiag = std::max(maxdiag, std::fabs(alpha));
    }
    const double thr = 2.220446049250313e-16 * n * maxdiag;
    int r = 0;
    for (int j = 0; j < rank; j++) if (std::fabs(A(j, j)) > thr) r = j + 1; else break;
    W = Mat(n, k, 0.0);
    for (int jj = 0; jj < k; jj++) {
        std::vector<double> z(n, 0.0);
        for (int i = r - 1; i >= 0; i--) {
            double s = B(i, jj);
            for (int t = i + 1; t < r; t++) s -= A(i, t) * z[t];
            z[i] = s / A(i, i);
        }
        for (int i = 0; i < n; i++) W(perm[i], jj) = z[i];
    }
}

// ---------------------------------------------------------------------------------------------
// LLE weights (trackdlo.cpp:92-159).  The op order below is mirrored exactly (no FMA
// contraction: build with -ffp-contract=off) by the device code so both produce identical bits.
// ---------------------------------------------------------------------------------------------
int lle_neighbours(int k, int M, int idx, int* out) {   // trackdlo.cpp:92-117
    int cnt = 0;
    if (idx - k < 0) {
        // UB: the reference reads up to idx+k unchecked (out of range when M <= idx+k); clip.
        for (int i = 0; i <= idx + k && i < M; i++) if (i != idx) out[cnt++] = i;
    } else if (idx + k >= M) {
        for (int i = idx - k; i <= M - 1; i++) if (i != idx) out[cnt++] = i;
    } else {
        for (int i = idx - k; i <= idx + k; i++) if (i != idx) out[cnt++] = i;
    }
    return cnt;
}

// partial-pivot LU of an nb x nb matrix stored row-major with stride 6; returns sign, fills piv
int lu6(double a[36], int nb, int piv[6]) {
    int sign = 1;
    for (int k = 0; k < nb; k++) {
        int p = k;
        double best = std::fabs(a[k * 6 + k]);
        for (int i = k + 1; i < nb; i++) {
            double v = std::fabs(a[i * 6 + k]);
            if (v > best) { best = v; p = i; }
        }
        piv[k] = p;
        if (p != k) {
            for (int j = 0; j < nb; j++) { double t = a[k * 6 + j]; a[k * 6 + j] = a[p * 6 + j]; a[p * 6 + j] = t; }
            sign = -sign;
        }
        double pv = a[k * 6 + k];
        if (pv != 0.0) {
            for (int i = k + 1; i < nb; i++) {
                double l = a[i * 6 + k] / pv;
                a[i * 6 + k] = l;
                for (int j = k + 1; j < nb; j++) a[i * 6 + j] = a[i * 6 + j] - l * a[k * 6 + j];
            }
        }
    }
    return sign;
}

// weights for one node; w[0..nb) ; returns nb
int lle_weights_node(const double* Y /*row-major M x 3*/, int M, int i, int* nbr, double* w) {
    int nb = lle_neighbours(3, M, i, nbr);                 // k/2 with k = 6 (trackdlo.cpp:122,236)
    double comp[6][3];
    for (int r = 0; r < nb; r++)
        for (int d = 0; d < 3; d++) comp[r][d] = Y[i * 3 + d] - Y[nbr[r] * 3 + d];
    double g[36] = {0}, lu[36];
    for (int a = 0; a < nb; a++)
        for (int b = 0; b < nb; b++)
            g[a * 6 + b] = (comp[a][0] * comp[b][0] + comp[a][1] * comp[b][1]) + comp[a][2] * comp[b][2];
    int piv[6];
    for (int t = 0; t < 36; t++) lu[t] = g[t];
    int sign = lu6(lu, nb, piv);
    double det = sign;
    for (int k = 0; k < nb; k++) det = det * lu[k * 6 + k];
    if (!(det != 0.0)) {                                    // trackdlo.cpp:136-144
        for (int k = 0; k < nb; k++) g[k * 6 + k] = g[k * 6 + k] + 0.00001;
        for (int t = 0; t < 36; t++) lu[t] = g[t];
        lu6(lu, nb, piv);
    }
    // explicit inverse column by column, then row sums (Gi_inv * 1) (trackdlo.cpp:150)
    double inv[36];
    for (int c = 0; c < nb; c++) {
        double b[6];
        for (int r = 0; r < nb; r++) b[r] = (r == c) ? 1.0 : 0.0;
        for (int k = 0; k < nb; k++) { int p = piv[k]; if (p != k) { double t = b[k]; b[k] = b[p]; b[p] = t; } }
        for (int r = 1; r < nb; r++) { double s = b[r]; for (int t = 0; t < r; t++) s = s - lu[r * 6 + t] * b[t]; b[r] = s; }
        for (int r = nb - 1; r >= 0; r--) {
            double s = b[r];
            for (int t = r + 1; t < nb; t++) s = s - lu[r * 6 + t] * b[t];
            b[r] = s / lu[r * 6 + r];
        }
        for (int r = 0; r < nb; r++) inv[r * 6 + c] = b[r];
    }
    double rs[6], tot = 0.0;
    for (int r = 0; r < nb; r++) {
        double s = 0.0;
        for (int c = 0; c < nb; c++) s = s + inv[r * 6 + c];
        rs[r] = s;
        tot = tot + s;
    }
    for (int r = 0; r < nb; r++) w[r] = rs[r] / tot;
    return nb;
}

// H = (I - L)^T (I - L), row-major M x M (trackdlo.cpp:236-237)
void lle_H(const double* Y, int M, std::vector<double>& H) {
    std::vector<double> E((size_t)M * M, 0.0);              // E = I - L
    for (int i = 0; i < M; i++) {
        int nbr[6]; double w[6];
        int nb = lle_weights_node(Y, M, i, nbr, w);
        E[(size_t)i * M + i] = 1.0;
        for (int r = 0; r < nb; r++) E[(size_t)i * M + nbr[r]] = 0.0 - w[r];
    }
    H.assign((size_t)M * M, 0.0);
    for (int a = 0; a < M; a++)
        for (int b = 0; b < M; b++) {
            double s = 0.0;
            for (int k = 0; k < M; k++) s = s + E[(size_t)k * M + a] * E[(size_t)k * M + b];
            H[(size_t)a * M + b] = s;
        }
}

struct CpdParams {
    double beta, lambda, lle_weight, mu, tol, alpha, k_vis, visibility_threshold;
    int max_iter, include_lle;
};

struct Trace {          // optional per-iteration capture; any pointer may be null
    double* P1;         // [max_iter][Nn]
    double* PX;         // [max_iter][Nn][3]
    double* Np;         // [max_iter]
    double* sigma2;     // [max_iter]   sigma2 after the update of that iteration
    double* W;          // [max_iter][Nn][3]
    double* Y;          // [max_iter][Nn][3] node positions after that iteration
    double* A;          // [Nn][Nn] row-major, first iteration only
    double* B;          // [Nn][3], first iteration only
};

// ---------------------------------------------------------------------------------------------
// cpd_lle (trackdlo.cpp:161-441), statement order kept.
// ---------------------------------------------------------------------------------------------
bool cpd_lle(const Mat& X_orig, Mat& Y, double& sigma2, const CpdParams& p,
             const std::vector<std::array<double, 4>>* priors_in, const std::vector<int>* vis_in,
             const double* H_override, Mat* W_out, int* iters_out, int64_t* kept_out, Trace* tr) {
    static const std::vector<std::array<double, 4>> no_priors;
    static const std::vector<int> no_vis;
    const auto& priors = priors_in ? *priors_in : no_priors;
    const auto& visible_nodes = vis_in ? *vis_in : no_vis;
    const int M = Y.r;
    const int D = 3;

    // prune X (trackdlo.cpp:177-195)
    Mat X_temp(X_orig.r, 3);
    int valid = 0;
    for (int i = 0; i < X_orig.r; i++) {
        double shortest = 100000;
        Vec3 xi = row3(X_orig, i);
        for (int j = 0; j < M; j++) {
            double dd = dist(row3(Y, j), xi);
            if (dd < shortest) shortest = dd;
        }
        if (shortest < 0.1) {
            X_temp(valid, 0) = xi.x; X_temp(valid, 1) = xi.y; X_temp(valid, 2) = xi.z;
            valid++;
        }
    }
    Mat X(valid, 3);
    for (int i = 0; i < valid; i++) for (int d = 0; d < 3; d++) X(i, d) = X_temp(i, d);
    const int N = valid;
    if (kept_out) *kept_out = N;

    bool converged = true;
    Mat Y_0 = Y;

    // arc-length coordinates from this call's Y_0 (trackdlo.cpp:216-223)
    std::vector<double> s(1, 0.0);
    double cur = 0;
    for (int i = 0; i < M - 1; i++) { cur += dist(row3(Y_0, i + 1), row3(Y_0, i)); s.push_back(cur); }

    // kernel (trackdlo.cpp:225-233); abs() is fabs
    Mat G(M, M);
    const double beta = p.beta;
    for (int i = 0; i < M; i++)
        for (int j = 0; j < M; j++) {
            double dd = std::fabs(s[i] - s[j]);
            G(i, j) = 1 / (2 * beta * 2 * beta) * std::exp(-std::sqrt(2.0) * dd / beta) * (2 * dd + std::sqrt(2.0) * beta);
        }

    // LLE matrix H (trackdlo.cpp:236-237): always built by the reference, used iff include_lle
    Mat H(M, M);
    if (p.include_lle) {
        if (H_override) {
            for (int i = 0; i < M; i++) for (int j = 0; j < M; j++) H(i, j) = H_override[(size_t)i * M + j];
        } else {
            std::vector<double> Yr((size_t)M * 3), Hr;
            for (int i = 0; i < M; i++) for (int d = 0; d < 3; d++) Yr[i * 3 + d] = Y_0(i, d);
            lle_H(Yr.data(), M, Hr);
            for (int i = 0; i < M; i++) for (int j = 0; j < M; j++) H(i, j) = Hr[(size_t)i * M + j];
        }
    }

    // priors -> J (diagonal 0/1), Y_extended (trackdlo.cpp:240-260)
    std::vector<double> Jd(M, 0.0);
    Mat Y_ext = Y_0;
    for (size_t i = 0; i < priors.size(); i++) {
        int index = (int)priors[i][0];
        if (index < 0 || index >= M) continue;           // UB: the reference would write out of range
        Jd[index] = 1.0;
        Y_ext(index, 0) = priors[i][1]; Y_ext(index, 1) = priors[i][2]; Y_ext(index, 2) = priors[i][3];
    }
    const bool have_priors = !priors.empty();

    // sigma2 init (trackdlo.cpp:263-273)
    Mat diff_xy(M, N);
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++) diff_xy(i, j) = sqdist(row3(Y_0, i), row3(X, j));
    if (sigma2 == 0) {
        double tot = 0;
        for (size_t t = 0; t < diff_xy.d.size(); t++) tot += diff_xy.d[t];
        sigma2 = tot / static_cast<double>((double)D * M * N);
    }

    Mat HG, HY0;
    if (p.include_lle) {            // iteration-invariant products (reference recomputes them each iteration)
        HG = Mat(M, M); HY0 = Mat(M, 3);
        for (int i = 0; i < M; i++) {
            for (int j = 0; j < M; j++) { double a = 0; for (int k = 0; k < M; k++) a += H(i, k) * G(k, j); HG(i, j) = a; }
            for (int d = 0; d < 3; d++) { double a = 0; for (int k = 0; k < M; k++) a += H(i, k) * Y_0(k, d); HY0(i, d) = a; }
        }
    }

    Mat P(M, N), geo(M, N), W(M, 3);
    std::vector<double> dmin(M), colsum(N), Pt1(N), P1(M);
    Mat PX(M, 3), T(M, 3), A(M, M), B(M, 3);
    int iters = 0;

    for (int it = 0; it < p.max_iter; it++) {
        iters = it + 1;
        // E-step 1 (trackdlo.cpp:278-296)
        for (int m = 0; m < M; m++) {
            double shortest = 10000;
            Vec3 ym = row3(Y, m);
            for (int n = 0; n < N; n++) {
                Vec3 xn = row3(X, n);
                diff_xy(m, n) = sqdist(ym, xn);
                double dd = dist(ym, xn);
                if (dd < shortest) shortest = dd;
            }
            if (shortest <= p.visibility_threshold) shortest = 0;
            dmin[m] = shortest;
        }
        // E-step 2 (trackdlo.cpp:298-301)
        for (size_t t = 0; t < P.d.size(); t++) P.d[t] = std::exp(-0.5 * diff_xy.d[t] / sigma2);
        double c = std::pow(2 * M_PI * sigma2, static_cast<double>(D) / 2) * p.mu / (1 - p.mu) * static_cast<double>(M) / N;
        for (int n = 0; n < N; n++) {
            double cs = 0;
            for (int m = 0; m < M; m++) cs += P(m, n);
            double den = cs + c;
            for (int m = 0; m < M; m++) P(m, n) = P(m, n) / den;
        }
        // E-step 3: geodesic distances (trackdlo.cpp:304-351)
        std::fill(geo.d.begin(), geo.d.end(), 0.0);
        for (int i = 0; i < N; i++) {
            int a = 0;
            double best = P(0, i);
            for (int m = 1; m < M; m++) if (P(m, i) > best) { best = P(m, i); a = m; }   // first max
            int q1 = a - 1; if (q1 == -1) q1 = 2;
            int q2 = a + 1; if (q2 == M) q2 = M - 3;
            Vec3 xi = row3(X, i);
            int b = (dist(row3(Y, q1), xi) < dist(row3(Y, q2), xi)) ? q1 : q2;
            geo(a, i) = sqdist(row3(Y, a), xi);
            geo(b, i) = sqdist(row3(Y, b), xi);
            int lo = a < b ? a : b, hi = a < b ? b : a;
            double dlo = dist(row3(Y, lo), xi), dhi = dist(row3(Y, hi), xi);
            for (int j = 0; j < lo; j++) geo(j, i) = std::pow(std::fabs(s[j] - s[lo]) + dlo, 2);
            for (int j = hi; j < M; j++) geo(j, i) = std::pow(std::fabs(s[j] - s[hi]) + dhi, 2);
        }
        // E-step 4 (trackdlo.cpp:354-383)
        for (size_t t = 0; t < P.d.size(); t++) P.d[t] = std::exp(-0.5 * geo.d[t] / sigma2);
        if ((int)visible_nodes.size() != M && !visible_nodes.empty() && p.k_vis != 0) {
            double total = 0;
            std::vector<double> pv(M);
            for (int m = 0; m < M; m++) { pv[m] = std::exp(-p.k_vis * dmin[m]); total += pv[m]; }
            for (int m = 0; m < M; m++) {
                double v = pv[m] / total;
                for (int n = 0; n < N; n++) P(m, n) = P(m, n) * v;
            }
            c = std::pow(2 * M_PI * sigma2, static_cast<double>(D) / 2) * p.mu / (1 - p.mu) / N;
        }
        for (int n = 0; n < N; n++) {
            double cs = 0;
            for (int m = 0; m < M; m++) cs += P(m, n);
            double den = cs + c;
            for (int m = 0; m < M; m++) P(m, n) = P(m, n) / den;
        }
        // reductions (trackdlo.cpp:386-389)
        double Np = 0;
        for (int n = 0; n < N; n++) { double a = 0; for (int m = 0; m < M; m++) a += P(m, n); Pt1[n] = a; }
        for (int m = 0; m < M; m++) {
            double a = 0, ax = 0, ay = 0, az = 0;
            for (int n = 0; n < N; n++) {
                double pv = P(m, n);
                a += pv; ax += pv * X(n, 0); ay += pv * X(n, 1); az += pv * X(n, 2);
            }
            P1[m] = a; PX(m, 0) = ax; PX(m, 1) = ay; PX(m, 2) = az;
            Np += a;
        }
        // M-step (trackdlo.cpp:392-413)
        const double ls = p.lambda * sigma2;
        for (int i = 0; i < M; i++) {
            for (int j = 0; j < M; j++) {
                double a = P1[i] * G(i, j) + (i == j ? ls : 0.0);
                if (p.include_lle) a += sigma2 * p.lle_weight * HG(i, j);
                if (have_priors) a += p.alpha * Jd[i] * G(i, j);
                A(i, j) = a;
            }
            for (int d = 0; d < 3; d++) {
                double b = PX(i, d) - P1[i] * Y_0(i, d);
                if (p.include_lle) b -= sigma2 * p.lle_weight * HY0(i, d);
                if (have_priors) b += p.alpha * (Y_ext(i, d) - Y_0(i, d));
                B(i, d) = b;
            }
        }
        if (tr && it == 0) {
            if (tr->A) for (int i = 0; i < M; i++) for (int j = 0; j < M; j++) tr->A[(size_t)i * M + j] = A(i, j);
            if (tr->B) for (int i = 0; i < M; i++) for (int d = 0; d < 3; d++) tr->B[i * 3 + d] = B(i, d);
        }
        colpiv_qr_solve(A, B, W);                                              // trackdlo.cpp:415
        // update (trackdlo.cpp:417-437)
        for (int i = 0; i < M; i++)
            for (int d = 0; d < 3; d++) {
                double a = 0;
                for (int k = 0; k < M; k++) a += G(i, k) * W(k, d);
                T(i, d) = Y_0(i, d) + a;
            }
        double trXPX = 0, trPXT = 0, trTPT = 0;
        for (int n = 0; n < N; n++) trXPX += Pt1[n] * (X(n, 0) * X(n, 0) + X(n, 1) * X(n, 1) + X(n, 2) * X(n, 2));
        for (int m = 0; m < M; m++) {
            trPXT += PX(m, 0) * T(m, 0) + PX(m, 1) * T(m, 1) + PX(m, 2) * T(m, 2);
            trTPT += P1[m] * (T(m, 0) * T(m, 0) + T(m, 1) * T(m, 1) + T(m, 2) * T(m, 2));
        }
        sigma2 = (trXPX - 2 * trPXT + trTPT) / (Np * D);

        double moved = 0;
        for (int m = 0; m < M; m++) moved += dist(row3(Y, m), row3(T, m));
        bool done = moved / M < p.tol;
        Y = T;
        if (tr) {
            if (tr->P1) for (int m = 0; m < M; m++) tr->P1[(size_t)it * M + m] = P1[m];
            if (tr->PX) for (int m = 0; m < M; m++) for (int d = 0; d < 3; d++) tr->PX[((size_t)it * M + m) * 3 + d] = PX(m, d);
            if (tr->Np) tr->Np[it] = Np;
            if (tr->sigma2) tr->sigma2[it] = sigma2;
            if (tr->W) for (int m = 0; m < M; m++) for (int d = 0; d < 3; d++) tr->W[((size_t)it * M + m) * 3 + d] = W(m, d);
            if (tr->Y) for (int m = 0; m < M; m++) for (int d = 0; d < 3; d++) tr->Y[((size_t)it * M + m) * 3 + d] = Y(m, d);
        }
        if (done) break;
        if (it == p.max_iter - 1) { converged = false; break; }
    }
    if (W_out) *W_out = W;
    if (iters_out) *iters_out = iters;
    return converged;
}

// ---------------------------------------------------------------------------------------------
// utils.cpp:172-241
// ---------------------------------------------------------------------------------------------
bool is_between(Vec3 x, Vec3 a, Vec3 b) {
    const double xs[3] = {x.x, x.y, x.z}, as[3] = {a.x, a.y, a.z}, bs[3] = {b.x, b.y, b.z};
    bool in_bound = true;
    for (int i = 0; i < 3; i++) {
        if (!(as[i] - 0.0001 <= xs[i] && xs[i] <= bs[i] + 0.0001) &&
            !(bs[i] - 0.0001 <= xs[i] && xs[i] <= as[i] + 0.0001)) in_bound = false;
    }
    return in_bound;
}

int line_sphere(Vec3 A, Vec3 B, Vec3 C, double radius, Vec3 out[2]) {
    double a = sqdist(A, B);
    double b = 2 * ((B.x - A.x) * (A.x - C.x) + (B.y - A.y) * (A.y - C.y) + (B.z - A.z) * (A.z - C.z));
    double c = sqdist(A, C) - std::pow(radius, 2);
    double delta = std::pow(b, 2) - 4 * a * c;
    int cnt = 0;
    if (delta < 0) return 0;
    if (delta > 0) {
        double d1 = (-b + std::sqrt(delta)) / (2 * a);
        double d2 = (-b - std::sqrt(delta)) / (2 * a);
        Vec3 p1 = {A.x + d1 * (B.x - A.x), A.y + d1 * (B.y - A.y), A.z + d1 * (B.z - A.z)};
        Vec3 p2 = {A.x + d2 * (B.x - A.x), A.y + d2 * (B.y - A.y), A.z + d2 * (B.z - A.z)};
        if (is_between(p1, A, B)) out[cnt++] = p1;
        if (is_between(p2, A, B)) out[cnt++] = p2;
    } else {
        double d1 = -b / (2 * a);
        Vec3 p1 = {A.x + d1 * (B.x - A.x), A.y + d1 * (B.y - A.y), A.z + d1 * (B.z - A.z)};
        if (is_between(p1, A, B)) out[cnt++] = p1;
    }
    return cnt;
}

typedef std::array<double, 4> Pair4;

// One pure-pursuit step over guide segments (shared by the six loops of trackdlo.cpp:617-894).
// Scans segments i = from, from+dir, ... while seg_ok(i); segment i joins guide[i] -> guide[i+dir].
// Returns the segment index that yielded (or -1) and writes the new centre.
template <class SegOk>
int pursue(const Mat& guide, int from, int dir, SegOk seg_ok, Vec3& centre, double look_ahead) {
    for (int i = from; seg_ok(i); i += dir) {
        Vec3 A = row3(guide, i), B = row3(guide, i + dir);
        Vec3 xs[2];
        int n = line_sphere(A, B, centre, look_ahead, xs);
        if (n == 0) continue;
        if (n == 1 && dist(xs[0], B) > dist(centre, B)) continue;
        Vec3 pick = xs[0];
        if (n == 2 && !(dist(xs[0], B) <= dist(xs[1], B))) pick = xs[1];
        centre = pick;
        return i;
    }
    return -1;
}

// traverse_euclidean (trackdlo.cpp:584-898)
std::vector<Pair4> traverse_euclidean(const std::vector<double>& geo, const Mat& guide,
                                      const std::vector<int>& vis, int alignment, int align_idx, int* err) {
    std::vector<Pair4> out;
    const int G = (int)geo.size();
    const int R = guide.r;
    const int V = (int)vis.size();
    auto emit = [&](double idx, Vec3 p) { out.push_back({idx, p.x, p.y, p.z}); };
    if (R == 1) { emit(vis[0], row3(guide, 0)); return out; }

    if (alignment == 0) {
        emit(vis[0], row3(guide, 0));
        int cs = 0;
        for (int i = 0; i < V; i++) { if (i == vis[i]) cs++; else break; }
        if (cs == 0) { if (err) *err |= 1; return out; }   // UB: size()-1 wraps in the reference
        int last = 0, k = 0;
        Vec3 centre = row3(guide, 0);
        while (last + 1 <= cs - 1 && k + 1 <= G - 1) {
            double look = std::fabs(geo[k + 1] - geo[k]);
            int got = pursue(guide, last, +1, [&](int i) { return i + 1 <= cs - 1; }, centre, look);
            if (got < 0) break;
            last = got;
            emit(k + 1, centre);
            k++;
        }
    } else if (alignment == 1) {
        emit(vis[V - 1], row3(guide, R - 1));
        int cs = 0;
        for (int i = 1; i <= V; i++) { if (vis[V - i] == G - i) cs++; else break; }
        int last = R - 1, k = G - 1;
        Vec3 centre = row3(guide, R - 1);
        const int lowest = R - cs;                         // R - cs >= 0 (cs <= V == R)
        while (last - 1 >= lowest && k - 1 >= 0) {
            double look = std::fabs(geo[k] - geo[k - 1]);
            int got = pursue(guide, last, -1, [&](int i) { return i >= lowest + 1; }, centre, look);
            if (got < 0) break;
            last = got;
            emit(k - 1, centre);
            k--;
        }
    } else {
        if (align_idx < 0 || align_idx >= V || align_idx >= R) { if (err) *err |= 2; return out; }
        emit(vis[align_idx], row3(guide, align_idx));
        // towards the tail (trackdlo.cpp:755-823)
        int cs2 = 1;
        for (int i = align_idx + 1; i < V; i++) { if (vis[i] - vis[i - 1] == 1) cs2++; else break; }
        int last = align_idx, k = vis[align_idx];
        Vec3 centre = row3(guide, align_idx);
        while (last + 1 <= align_idx + cs2 - 1 && k + 1 <= G - 1) {
            double look = std::fabs(geo[k + 1] - geo[k]);
            int got = pursue(guide, last, +1, [&](int i) { return i + 1 <= align_idx + cs2 - 1; }, centre, look);
            if (got < 0) break;
            last = got;
            emit(k + 1, centre);
            k++;
        }
        // towards the head (trackdlo.cpp:826-894).  The reference's run loop increments i while
        // testing i >= 0 (trackdlo.cpp:828), i.e. it walks *up* from align_idx-1.
        // UB: it can read visible_nodes[size]; we stop at the end of the vector instead.
        int cs1 = 1;
        if (align_idx - 1 >= 0)
            for (int i = align_idx - 1; i >= 0 && i + 1 < V; i++) { if (vis[i + 1] - vis[i] == 1) cs1++; else break; }
        last = align_idx; k = vis[align_idx];
        centre = row3(guide, align_idx);
        // `last-1 >= align_idx - cs1.size()` is evaluated in unsigned 64-bit arithmetic (trackdlo.cpp:842)
        auto cond = [&](int l) {
            return (uint64_t)(int64_t)(l - 1) >= (uint64_t)(int64_t)align_idx - (uint64_t)cs1;
        };
        while (cond(last) && k - 1 >= 0) {
            double look = std::fabs(geo[k] - geo[k - 1]);
            int got = pursue(guide, last, -1, [&](int i) { return i - 1 >= 0; }, centre, look);
            if (got < 0) break;
            last = got;
            emit(k - 1, centre);
            k--;
        }
    }
    return out;
}

struct TrackParams {
    double visibility_threshold, beta, lambda, alpha, k_vis, mu, tol, beta_pre_proc, lambda_pre_proc, lle_weight;
    int max_iter;
};

// tracking_step (trackdlo.cpp:900-999)
int tracking_step(const Mat& X, Mat& Y_, double& sigma2_, const std::vector<double>& geodesic_coord,
                  const std::vector<int>& vis, const std::vector<int>& vis_ext, const TrackParams& tp,
                  const double* H_override, Mat& guide, std::vector<Pair4>& priors, int iters[2], int conv[2], int* state_out) {
    const int Nn = Y_.r;
    const int V = (int)vis_ext.size();
    int err = 0;
    guide = Mat(V, 3);
    if (V != Nn) { for (int i = 0; i < V; i++) for (int d = 0; d < 3; d++) guide(i, d) = Y_(vis_ext[i], d); }
    else guide = Y_;

    double sigma2_pre = sigma2_;
    CpdParams pre = {tp.beta_pre_proc, tp.lambda_pre_proc, tp.lle_weight, tp.mu, tp.tol, 0.0, 0.0, 0.01, tp.max_iter, 1};
    conv[0] = cpd_lle(X, guide, sigma2_pre, pre, nullptr, nullptr, H_override, nullptr, &iters[0], nullptr, nullptr);

    priors.clear();
    int state;
    if (V == Nn) {
        state = 0;
        std::vector<Pair4> v1 = traverse_euclidean(geodesic_coord, guide, vis_ext, 0, -1, &err);
        std::vector<Pair4> v2 = traverse_euclidean(geodesic_coord, guide, vis_ext, 1, -1, &err);
        std::reverse(v2.begin(), v2.end());
        const int n1 = (int)v1.size(), n2 = (int)v2.size();
        for (int i = 0; i < Nn; i++) {
            const int j2 = i - (Nn - n2);
            if (i < v2[0][0] && i < n1) priors.push_back(v1[i]);
            else if (i > v1[n1 - 1][0] && j2 >= 0 && j2 < n2) priors.push_back(v2[j2]);
            else if (i < n1 && j2 >= 0 && j2 < n2) {
                Pair4 a;
                for (int t = 0; t < 4; t++) a[t] = (v1[i][t] + v2[j2][t]) / 2.0;
                priors.push_back(a);
            } else err |= 4;                                 // UB: the reference indexes out of range here
        }
    } else if (vis_ext[0] == 0 && vis_ext[V - 1] == Nn - 1) {
        state = 1;
        priors = traverse_euclidean(geodesic_coord, guide, vis_ext, 0, -1, &err);
        std::vector<Pair4> v2 = traverse_euclidean(geodesic_coord, guide, vis_ext, 1, -1, &err);
        priors.insert(priors.end(), v2.begin(), v2.end());
    } else if (vis_ext[0] == 0) {
        state = 2;
        priors = traverse_euclidean(geodesic_coord, guide, vis_ext, 0, -1, &err);
    } else if (vis_ext[V - 1] == Nn - 1) {
        state = 3;
        priors = traverse_euclidean(geodesic_coord, guide, vis_ext, 1, -1, &err);
    } else {
        state = 4;
        int align = -1;
        double moved = 999999;
        for (int i = 0; i < (int)vis.size(); i++) {
            if (i >= guide.r) { err |= 8; break; }          // UB: guide indexed by positions of visible_nodes
            double dd = dist(row3(Y_, vis[i]), row3(guide, i));
            if (dd < moved) { moved = dd; align = i; }
        }
        priors = traverse_euclidean(geodesic_coord, guide, vis_ext, 2, align, &err);
    }
    if (state_out) *state_out = state;

    CpdParams mainp = {tp.beta, tp.lambda, tp.lle_weight, tp.mu, tp.tol, tp.alpha, tp.k_vis, tp.visibility_threshold, tp.max_iter, 0};
    conv[1] = cpd_lle(X, Y_, sigma2_, mainp, &priors, &vis_ext, nullptr, nullptr, &iters[1], nullptr, nullptr);
    return err;
}

Mat from_rows(const double* p, int64_t n) {
    Mat m((int)n, 3);
    for (int64_t i = 0; i < n; i++) for (int d = 0; d < 3; d++) m((int)i, d) = p[i * 3 + d];
    return m;
}
void to_rows(const Mat& m, double* p) {
    for (int i = 0; i < m.r; i++) for (int d = 0; d < m.c; d++) p[(size_t)i * m.c + d] = m(i, d);
}

}  // namespace

extern "C" {

struct oracle_cpd_params {
    double beta, lambda, lle_weight, mu, tol, alpha, k_vis, visibility_threshold;
    int32_t max_iter, include_lle;
};
struct oracle_trace { double *P1, *PX, *Np, *sigma2, *W, *Y, *A, *B; };
struct oracle_track_params {
    double visibility_threshold, beta, lambda, alpha, k_vis, mu, tol, beta_pre_proc, lambda_pre_proc, lle_weight;
    int32_t max_iter, pad_;
};

// returns 1 if converged else 0
int oracle_cpd_lle(const double* X, int64_t n_points, double* Y, int32_t n_nodes, double* sigma2,
                   const oracle_cpd_params* prm, const double* priors, int32_t n_priors,
                   const int32_t* vis, int32_t n_vis, const double* H_override,
                   double* W_out, int32_t* iters_out, int64_t* kept_out, const oracle_trace* trace) {
    Mat Xm = from_rows(X, n_points), Ym = from_rows(Y, n_nodes);
    CpdParams p = {prm->beta, prm->lambda, prm->lle_weight, prm->mu, prm->tol, prm->alpha, prm->k_vis,
                   prm->visibility_threshold, prm->max_iter, prm->include_lle};
    std::vector<Pair4> pr(n_priors);
    for (int i = 0; i < n_priors; i++) for (int t = 0; t < 4; t++) pr[i][t] = priors[i * 4 + t];
    std::vector<int> v(vis, vis + (vis ? n_vis : 0));
    Trace tr{};
    if (trace) tr = {trace->P1, trace->PX, trace->Np, trace->sigma2, trace->W, trace->Y, trace->A, trace->B};
    Mat W;
    int it = 0;
    bool conv = cpd_lle(Xm, Ym, *sigma2, p, &pr, &v, H_override, &W, &it, kept_out, trace ? &tr : nullptr);
    to_rows(Ym, Y);
    if (W_out && W.r == n_nodes) to_rows(W, W_out);
    if (iters_out) *iters_out = it;
    return conv ? 1 : 0;
}

void oracle_lle_H(const double* Y, int32_t n_nodes, double* H_out) {
    std::vector<double> H;
    lle_H(Y, n_nodes, H);
    std::memcpy(H_out, H.data(), H.size() * sizeof(double));
}

// returns the UB/err bitmask (0 = clean)
int oracle_tracking_step(const double* X, int64_t n_points, double* Y, int32_t n_nodes, double* sigma2,
                         const double* geodesic_coord, const int32_t* vis, int32_t n_vis,
                         const int32_t* vis_ext, int32_t n_vis_ext, const oracle_track_params* tp,
                         const double* H_override, double* guide_out, double* priors_out, int32_t* n_priors_out,
                         int32_t* iters_out /*[2]*/, int32_t* converged_out /*[2]*/, int32_t* state_out) {
    Mat Xm = from_rows(X, n_points), Ym = from_rows(Y, n_nodes);
    std::vector<double> geo(geodesic_coord, geodesic_coord + n_nodes);
    std::vector<int> v(vis, vis + n_vis), ve(vis_ext, vis_ext + n_vis_ext);
    TrackParams t = {tp->visibility_threshold, tp->beta, tp->lambda, tp->alpha, tp->k_vis, tp->mu, tp->tol,
                     tp->beta_pre_proc, tp->lambda_pre_proc, tp->lle_weight, tp->max_iter};
    Mat guide;
    std::vector<Pair4> priors;
    int iters[2] = {0, 0}, conv[2] = {0, 0}, state = -1;
    int err = tracking_step(Xm, Ym, *sigma2, geo, v, ve, t, H_override, guide, priors, iters, conv, &state);
    to_rows(Ym, Y);
    if (guide_out) to_rows(guide, guide_out);
    if (priors_out) for (size_t i = 0; i < priors.size(); i++) for (int k = 0; k < 4; k++) priors_out[i * 4 + k] = priors[i][k];
    if (n_priors_out) *n_priors_out = (int)priors.size();
    if (iters_out) { iters_out[0] = iters[0]; iters_out[1] = iters[1]; }
    if (converged_out) { converged_out[0] = conv[0]; converged_out[1] = conv[1]; }
    if (state_out) *state_out = state;
    return err;
}

int oracle_traverse_euclidean(const double* geodesic_coord, int32_t n_geo, const double* guide, int32_t n_guide,
                              const int32_t* vis, int32_t n_vis, int32_t alignment, int32_t align_idx,
                              double* pairs_out, int32_t* n_pairs_out) {
    std::vector<double> geo(geodesic_coord, geodesic_coord + n_geo);
    Mat g = from_rows(guide, n_guide);
    std::vector<int> v(vis, vis + n_vis);
    int err = 0;
    std::vector<Pair4> out = traverse_euclidean(geo, g, v, alignment, align_idx, &err);
    for (size_t i = 0; i < out.size(); i++) for (int k = 0; k < 4; k++) pairs_out[i * 4 + k] = out[i][k];
    *n_pairs_out = (int)out.size();
    return err;
}

// Visibility front-end that feeds tracking_step: trackdlo/src/trackdlo_node.cpp:254-277 (shortest
// node-to-point distances, initial value 100000 as in the reference) and :346-360 (visible_nodes sorted ascending,
// visible_nodes_extended by the d_vis rule on converted_node_coord).  The self-occlusion raster (:280-343) is NOT
// part of this restatement: every node counts as not self-occluded.  An empty visible list gives an empty extended
// list (UB: the reference evaluates visible_nodes.size()-1 on an empty vector).
// Returns the number of visible nodes; *n_ext_out the number of extended ones.
int oracle_visibility(const double* X, int64_t Mp, const double* Y, int32_t Nn, const double* node_coord,
                      double visibility_threshold, double d_vis, double* dmin_out, int32_t* vis_out, int32_t* ext_out,
                      int32_t* n_ext_out) {
    int nv = 0;
    for (int m = 0; m < Nn; m++) {
        double shortest = 100000;
        for (int64_t n = 0; n < Mp; n++) {
            const double dx = Y[3 * m] - X[3 * n], dy = Y[3 * m + 1] - X[3 * n + 1], dz = Y[3 * m + 2] - X[3 * n + 2];
            const double dist = std::sqrt(dx * dx + dy * dy + dz * dz);
            if (dist < shortest) shortest = dist;
        }
        if (dmin_out) dmin_out[m] = shortest;
        if (shortest <= visibility_threshold) vis_out[nv++] = m;      // ascending m == std::sort of the reference's list
    }
    int ne = 0;
    if (nv > 0) {
        for (int i = 0; i + 1 < nv; i++) {
            ext_out[ne++] = vis_out[i];
            if (std::fabs(node_coord[vis_out[i + 1]] - node_coord[vis_out[i]]) <= d_vis)
                for (int j = 1; j < vis_out[i + 1] - vis_out[i]; j++) ext_out[ne++] = vis_out[i] + j;
        }
        ext_out[ne++] = vis_out[nv - 1];
    }
    *n_ext_out = ne;
    return nv;
}

// Evaluator error metric: trackdlo/src/evaluator.cpp:233-283 (calc_min_distance, get_piecewise_error)
// and :333-341 (compute_error), helpers utils.cpp:477-489.  Mean over the nodes of one polyline of the distance to the
// nearest segment of the other, symmetrised.
static double seg_point_distance(const double* A, const double* B, const double* E) {
    const double AB[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]}, AE[3] = {E[0] - A[0], E[1] - A[1], E[2] - A[2]};
    const double cx = AE[1] * AB[2] - AE[2] * AB[1], cy = -(AE[0] * AB[2] - AE[2] * AB[0]), cz = AE[0] * AB[1] - AE[1] * AB[0];
    const double abab = AB[0] * AB[0] + AB[1] * AB[1] + AB[2] * AB[2];
    const double aeab = AE[0] * AB[0] + AE[1] * AB[1] + AE[2] * AB[2];
    double distance = std::sqrt(cx * cx + cy * cy + cz * cz) / std::sqrt(abab);
    const double P[3] = {A[0] + AB[0] * aeab / abab, A[1] + AB[1] * aeab / abab, A[2] + AB[2] * aeab / abab};
    const double AP[3] = {P[0] - A[0], P[1] - A[1], P[2] - A[2]};
    const double apab = AP[0] * AB[0] + AP[1] * AB[1] + AP[2] * AB[2];
    if (apab < 0 || apab > abab) {
        const double BE[3] = {E[0] - B[0], E[1] - B[1], E[2] - B[2]};
        const double dAE = std::sqrt(AE[0] * AE[0] + AE[1] * AE[1] + AE[2] * AE[2]);
        const double dBE = std::sqrt(BE[0] * BE[0] + BE[1] * BE[1] + BE[2] * BE[2]);
        distance = dAE > dBE ? dBE : dAE;
    }
    return distance;
}
static double piecewise_error(const double* Yt, int nt, const double* Yr, int nr) {
    double total = 0.0;
    for (int idx = 0; idx < nt; idx++) {
        double dist = -1;
        for (int i = 0; i < nr - 1; i++) {
            const double di = seg_point_distance(Yr + 3 * i, Yr + 3 * (i + 1), Yt + 3 * idx);
            if (dist == -1 || di < dist) dist = di;
        }
        total += dist;
    }
    return total / nt;
}
double oracle_tracking_error(const double* Y_track, int32_t n_track, const double* Y_true, int32_t n_true) {
    return (piecewise_error(Y_track, n_track, Y_true, n_true) + piecewise_error(Y_true, n_true, Y_track, n_track)) / 2;
}

}  // extern "C"
