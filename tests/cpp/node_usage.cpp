// Uses `class trackdlo` the way the reference's only caller does (trackdlo_node.cpp:54 global instance, :131 copy-assignment
// of the 12-argument constructor, :136-143 initialize_nodes / initialize_geodesic_coord, :366-369 tracking_step + the three
// getters), through "trackdlo.h" -- which tests/test_dropin.py points either at the REAL reference header with the
// INTEGRATION.md recipe applied, or at tests/cpp/reference_header_shim.hpp.  Reads a frame from a raw binary file and
// writes the results (same format as adapter_run.cpp).
// File in : int64 {Nn, Mp, n_vis, n_ext}, double params[12], Y[Nn*3] (row-major), rest[Nn], X[Mp*3], int32 vis[], ext[]
// File out: double Y[Nn*3], sigma2, guide[n_ext*3], n_priors, priors[n_priors*4]
#include "trackdlo.h"
#include <cstdio>
#include <cstdlib>

trackdlo tracker;                                   // trackdlo_node.cpp:54
MatrixXd Y, guide_nodes;
std::vector<MatrixXd> priors;
std::vector<double> converted_node_coord = {0.0};
std::vector<int> visible_nodes, visible_nodes_extended;

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 3;
    int64_t hdr[4];
    double prm[12];
    if (std::fread(hdr, 8, 4, f) != 4 || std::fread(prm, 8, 12, f) != 12) return 4;
    const int Nn = (int)hdr[0]; const long Mp = (long)hdr[1]; const int nv = (int)hdr[2], ne = (int)hdr[3];
    std::vector<double> Yr((size_t)Nn * 3), rest(Nn), Xr((size_t)Mp * 3);
    std::vector<int32_t> vis(nv), ext(ne);
    if (std::fread(Yr.data(), 8, Yr.size(), f) != Yr.size() || std::fread(rest.data(), 8, Nn, f) != (size_t)Nn ||
        std::fread(Xr.data(), 8, Xr.size(), f) != Xr.size() || std::fread(vis.data(), 4, nv, f) != (size_t)nv ||
        std::fread(ext.data(), 4, ne, f) != (size_t)ne) return 5;
    std::fclose(f);
    MatrixXd init_nodes = MatrixXd::Zero(Nn, 3), X = MatrixXd::Zero(Mp, 3), proj_matrix = MatrixXd::Zero(3, 4);
    for (int i = 0; i < Nn; i++) for (int d = 0; d < 3; d++) init_nodes(i, d) = Yr[(size_t)i * 3 + d];
    for (long i = 0; i < Mp; i++) for (int d = 0; d < 3; d++) X(i, d) = Xr[(size_t)i * 3 + d];
    const double visibility_threshold = prm[0], beta = prm[1], lambda = prm[2], alpha = prm[3], k_vis = prm[4], mu = prm[5];
    const int max_iter = (int)prm[6];
    const double tol = prm[7], beta_pre_proc = prm[8], lambda_pre_proc = prm[9], lle_weight = prm[10];

    tracker = trackdlo(init_nodes.rows(), visibility_threshold, beta, lambda, alpha, k_vis, mu, max_iter, tol, beta_pre_proc, lambda_pre_proc, lle_weight);   // :131
    for (int i = 1; i < Nn; i++) converted_node_coord.push_back(rest[i]);                // :136-140 (arc lengths come with the input here)
    tracker.initialize_nodes(init_nodes);                                                // :142
    tracker.initialize_geodesic_coord(converted_node_coord);                             // :143
    visible_nodes.assign(vis.begin(), vis.end());
    visible_nodes_extended.assign(ext.begin(), ext.end());
    tracker.tracking_step(X, visible_nodes, visible_nodes_extended, proj_matrix, 720, 1280);   // :366
    Y = tracker.get_tracking_result();                                                   // :367
    guide_nodes = tracker.get_guide_nodes();                                             // :368
    priors = tracker.get_correspondence_pairs();                                         // :369

    FILE* o = std::fopen(argv[2], "wb");
    if (!o) return 6;
    for (int i = 0; i < Nn; i++) for (int d = 0; d < 3; d++) { double v = Y(i, d); std::fwrite(&v, 8, 1, o); }
    double s2 = tracker.get_sigma2(); std::fwrite(&s2, 8, 1, o);
    for (long i = 0; i < guide_nodes.rows(); i++) for (int d = 0; d < 3; d++) { double v = guide_nodes(i, d); std::fwrite(&v, 8, 1, o); }
    double np = (double)priors.size(); std::fwrite(&np, 8, 1, o);
    for (size_t k = 0; k < priors.size(); k++) for (int t = 0; t < 4; t++) { double v = priors[k](0, t); std::fwrite(&v, 8, 1, o); }
    std::fclose(o);
    return 0;
}
