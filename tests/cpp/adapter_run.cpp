// Runs the drop-in `class trackdlo` (include/trackdlo_adapter.hpp) on a frame read from a raw
// binary file and writes the results; driven by tests/test_gpu_parity.py::test_adapter_class_*.
// File in : int64 {Nn, Mp, n_vis, n_ext}, double params[12], Y[Nn*3] (row-major), rest[Nn], X[Mp*3], int32 vis[], ext[]
// File out: double Y[Nn*3], sigma2, guide[n_ext*3], n_priors, priors[n_priors*4]
#define TRACKDLO_ADAPTER_MATRIX_HEADER "matrix_stub.hpp"
#include "trackdlo_adapter.hpp"
#include <cstdio>
#include <cstdlib>

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 3;
    int64_t hdr[4];
    double prm[12];
    if (std::fread(hdr, 8, 4, f) != 4 || std::fread(prm, 8, 12, f) != 12) return 4;
    const int Nn = (int)hdr[0]; const long Mp = (long)hdr[1]; const int nv = (int)hdr[2], ne = (int)hdr[3];
    std::vector<double> Y((size_t)Nn * 3), rest(Nn), X((size_t)Mp * 3);
    std::vector<int32_t> vis(nv), ext(ne);
    if (std::fread(Y.data(), 8, Y.size(), f) != Y.size() || std::fread(rest.data(), 8, Nn, f) != (size_t)Nn ||
        std::fread(X.data(), 8, X.size(), f) != X.size() || std::fread(vis.data(), 4, nv, f) != (size_t)nv ||
        std::fread(ext.data(), 4, ne, f) != (size_t)ne) return 5;
    std::fclose(f);
    MatrixXd Ym(Nn, 3), Xm(Mp, 3), proj(3, 4);
    for (int i = 0; i < Nn; i++) for (int d = 0; d < 3; d++) Ym(i, d) = Y[(size_t)i * 3 + d];
    for (long i = 0; i < Mp; i++) for (int d = 0; d < 3; d++) Xm(i, d) = X[(size_t)i * 3 + d];
    trackdlo tracker;
    // trackdlo(num_of_nodes, visibility_threshold, beta, lambda, alpha, k_vis, mu, max_iter, tol, beta_pre, lambda_pre, lle_weight)
    tracker = trackdlo(Nn, prm[0], prm[1], prm[2], prm[3], prm[4], prm[5], (int)prm[6], prm[7], prm[8], prm[9], prm[10]);
    tracker.initialize_nodes(Ym);
    tracker.initialize_geodesic_coord(rest);
    tracker.tracking_step(Xm, std::vector<int>(vis.begin(), vis.end()), std::vector<int>(ext.begin(), ext.end()), proj, 720, 1280);
    MatrixXd out = tracker.get_tracking_result(), guide = tracker.get_guide_nodes();
    std::vector<MatrixXd> pri = tracker.get_correspondence_pairs();
    FILE* o = std::fopen(argv[2], "wb");
    if (!o) return 6;
    for (int i = 0; i < Nn; i++) for (int d = 0; d < 3; d++) { double v = out(i, d); std::fwrite(&v, 8, 1, o); }
    double s2 = tracker.get_sigma2(); std::fwrite(&s2, 8, 1, o);
    for (long i = 0; i < guide.rows(); i++) for (int d = 0; d < 3; d++) { double v = guide(i, d); std::fwrite(&v, 8, 1, o); }
    double np = (double)pri.size(); std::fwrite(&np, 8, 1, o);
    for (size_t k = 0; k < pri.size(); k++) for (int t = 0; t < 4; t++) { double v = pri[k](0, t); std::fwrite(&v, 8, 1, o); }
    std::fclose(o);
    return tracker.last_status() & ~(TDLO_ST_NOT_CONVERGED | TDLO_ST_PRE_NOT_CONVERGED);
}
