// Drop-in for /root/reference/include/gp_regression/cov_functions.h and kernels/*.hpp.
//
// The three covariance classes keep the reference's public surface — compute / computediff /
// computediffdiff, the same constructors and defaults, public sigma_/length_ on Gaussian and Laplace —
// so caller code such as std::make_shared<ThinPlate>(2.0) (src/gp_node.cpp:919) compiles unchanged.
// On the hot path they are only parameter carriers: descriptor() hands (kind, p0, p1) to the CUDA
// kernels, which evaluate the same formulas on the device (csrc/gpr_common.cuh).
//   ThinPlate  k(d) = 2d^3 - 3Rd^2 + R^3,  k~(d) = -6(R - d)          kernels/thin_plate.hpp:14,:19
//   Gaussian   k(d) = sigma^2 exp(-d/length^2), k~ = -k/length^2      kernels/gaussian.hpp:17-18,:24-25
//   Laplace    k(d) = 2 sigma exp(-d/length),   k~ = -k/length        kernels/laplace.hpp:39-40,:46-47
// (d is the Euclidean distance, not its square: SURVEY F3.  computediffdiff is 0 in the reference.)
#pragma once
#include <cmath>
#include "../gpr_c_api.h"

namespace gp_regression {

class ThinPlate {
public:
    ThinPlate() : ThinPlate(1.0) {}
    ThinPlate(double R) : radius_(R), cube_(R * R * R) {}
    double compute(double d) const { return 2 * d * d * d - 3 * radius_ * d * d + cube_; }
    double computediff(double d) const { return -6 * (radius_ - d); }
    double computediffdiff(double) const { return 0; }
    double R() const { return radius_; }
    gpr_kernel_t descriptor() const { return gpr_kernel_t{0, radius_, 0.0}; }
private:
    double radius_, cube_;
};

class Gaussian {
public:
    const double sigma_;
    const double length_;
    Gaussian() : Gaussian(1.0, 1.0) {}
    Gaussian(double sigma, double length)
        : sigma_(sigma), length_(length), amp_(sigma * sigma), rate_(1.0 / (length * length)) {}
    double compute(const double& d) const { return amp_ * std::exp(-1 * d * rate_); }
    double computediff(const double& d) const { return -1 * rate_ * compute(d); }
    double computediffdiff(const double&) const { return 0.0; }
    gpr_kernel_t descriptor() const { return gpr_kernel_t{1, sigma_, length_}; }
private:
    double amp_, rate_;
};

class Laplace {
public:
    const double sigma_;
    const double length_;
    Laplace() : Laplace(1.0, 1.0) {}
    Laplace(double sigma, double length) : sigma_(sigma), length_(length), rate_(1.0 / length) {}
    double compute(const double& d) const { return 2 * sigma_ * std::exp(-1 * d * rate_); }
    double computediff(const double& d) const { return -1 * rate_ * compute(d); }
    double computediffdiff(const double&) const { return 0.0; }
    gpr_kernel_t descriptor() const { return gpr_kernel_t{2, sigma_, length_}; }
private:
    double rate_;
};

}  // namespace gp_regression
