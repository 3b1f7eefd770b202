// mini_eigen.h — used only when Eigen is not installed (it is absent from the build image).
// The hot path needs no linear algebra on the host; callers only need containers with the handful of
// accessors they use on the reference's outputs: rows/cols/size/data, (i,j), (i), row(i), resize,
// resizeLike, and Vector3d.  With Eigen present the real types are used instead (gp_regressor.hpp).
#pragma once
#include <cmath>
#include <cstddef>
#include <vector>

namespace Eigen {

typedef std::ptrdiff_t Index;

struct Vector3d {
    double v[3];
    Vector3d() : v{0, 0, 0} {}
    Vector3d(double x, double y, double z) : v{x, y, z} {}
    double& operator()(Index i) { return v[i]; }
    double operator()(Index i) const { return v[i]; }
    double& operator[](Index i) { return v[i]; }
    double operator[](Index i) const { return v[i]; }
    double norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
    double dot(const Vector3d& o) const { return v[0] * o.v[0] + v[1] * o.v[1] + v[2] * o.v[2]; }
    Vector3d cross(const Vector3d& o) const {
        return Vector3d(v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]);
    }
    void normalize() { double n = norm(); if (n > 0) { v[0] /= n; v[1] /= n; v[2] /= n; } }
    Vector3d normalized() const { Vector3d t(*this); t.normalize(); return t; }
};

class MatrixXd {                       // column-major, like Eigen's default
public:
    MatrixXd() : r_(0), c_(0) {}
    MatrixXd(Index r, Index c) : r_(r), c_(c), d_((size_t)(r * c), 0.0) {}
    Index rows() const { return r_; }
    Index cols() const { return c_; }
    Index size() const { return r_ * c_; }
    double* data() { return d_.data(); }
    const double* data() const { return d_.data(); }
    void resize(Index r, Index c) { r_ = r; c_ = c; d_.assign((size_t)(r * c), 0.0); }
    void resize(Index n) { resize(n, 1); }
    template <class M> void resizeLike(const M& o) { resize(o.rows(), o.cols()); }
    double& operator()(Index i, Index j) { return d_[(size_t)(j * r_ + i)]; }
    double operator()(Index i, Index j) const { return d_[(size_t)(j * r_ + i)]; }
    double& operator()(Index i) { return d_[(size_t)i]; }
    double operator()(Index i) const { return d_[(size_t)i]; }
    Vector3d row(Index i) const { return Vector3d((*this)(i, 0), (*this)(i, 1), (*this)(i, 2)); }
private:
    Index r_, c_;
    std::vector<double> d_;
};

typedef MatrixXd VectorXd;

}  // namespace Eigen
