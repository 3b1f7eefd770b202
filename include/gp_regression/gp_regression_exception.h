// Drop-in for /root/reference/include/gp_regression/gp_regression_exception.h:9-17.
// Same name, namespace, constructor and what(); it is a std::exception, so existing catch sites work.
// (The reference forgets to include <string>; this header does.)
#pragma once
#include <stdexcept>
#include <string>

namespace gp_regression {

class GPRegressionException : public std::runtime_error {
public:
    explicit GPRegressionException(const std::string& message, int status = 1, long long pivot = 0)
        : std::runtime_error(message), status_(status), pivot_(pivot) {}
    int status() const noexcept { return status_; }        // GPR_ERR_* of include/gpr_c_api.h
    long long pivot() const noexcept { return pivot_; }    // 1-based failing pivot for GPR_ERR_NOT_SPD
private:
    int status_;
    long long pivot_;
};

}  // namespace gp_regression
