// Compile-only check: the adapter exposes the reference's `class trackdlo` surface
// (trackdlo/include/trackdlo.h:53-103) and is default-constructible / copy-assignable the way
// trackdlo_node.cpp:54,131 uses it.
#define TRACKDLO_ADAPTER_MATRIX_HEADER "matrix_stub.hpp"
#include "trackdlo_adapter.hpp"

trackdlo tracker;   // global default-constructed instance (trackdlo_node.cpp:54)

int adapter_surface_check() {
    tracker = trackdlo(45, 0.008, 0.35, 50000, 3, 50, 0.1, 50, 0.0002, 3.0, 1.0, 10.0);   // trackdlo_node.cpp:131
    trackdlo small(10);
    MatrixXd Y = MatrixXd::Zero(45, 3), X = MatrixXd::Zero(100, 3), proj = MatrixXd::Zero(3, 4);
    tracker.initialize_nodes(Y);
    tracker.initialize_geodesic_coord(std::vector<double>(45, 0.0));
    tracker.set_sigma2(0.0);
    std::vector<int> vis = {0, 1, 2}, ext = {0, 1, 2};
    tracker.tracking_step(X, vis, ext, proj, 720, 1280);
    MatrixXd out = tracker.get_tracking_result();
    MatrixXd guide = tracker.get_guide_nodes();
    std::vector<MatrixXd> priors = tracker.get_correspondence_pairs();
    double s2 = tracker.get_sigma2();
    bool ok = tracker.cpd_lle(X, Y, s2, 0.35, 50000, 10.0, 0.1);
    ok = tracker.cpd_lle(X, Y, s2, 0.35, 50000, 10.0, 0.1, 50, 0.0002, false, priors, 3.0, vis, 50.0, 0.008) && ok;
    return (int)out.rows() + (int)guide.rows() + (int)priors.size() + (ok ? 1 : 0);
}
