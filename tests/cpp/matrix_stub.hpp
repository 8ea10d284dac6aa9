// Minimal column-major stand-in for Eigen::MatrixXd (Eigen is not installed in this image).
// Only what include/trackdlo_adapter.hpp touches: rows(), cols(), data(), operator(), Zero().
#pragma once
#include <cstddef>
#include <vector>
namespace Eigen {
class MatrixXd {
public:
    MatrixXd() {}
    MatrixXd(long r, long c) : r_(r), c_(c), d_((size_t)r * c, 0.0) {}
    static MatrixXd Zero(long r, long c) { return MatrixXd(r, c); }
    long rows() const { return r_; }
    long cols() const { return c_; }
    double* data() { return d_.data(); }
    const double* data() const { return d_.data(); }
    double& operator()(long i, long j) { return d_[(size_t)i + (size_t)j * r_]; }
    double operator()(long i, long j) const { return d_[(size_t)i + (size_t)j * r_]; }
private:
    long r_ = 0, c_ = 0;
    std::vector<double> d_;
};
}  // namespace Eigen
