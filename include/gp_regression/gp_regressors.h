// Drop-in for /root/reference/include/gp_regression/gp_regressors.h:11-18: the concrete regressor
// types (default-constructed kernels) and their Ptr/ConstPtr aliases.
#pragma once
#include <memory>
#include "gp_regressor.hpp"

namespace gp_regression {

#define GPR_DECLARE_REGRESSOR(Name, Cov)                                     \
    class Name : public GPRegressor<Cov> {                                   \
    public:                                                                  \
        typedef std::shared_ptr<Name> Ptr;                                   \
        typedef std::shared_ptr<const Name> ConstPtr;                        \
    }

GPR_DECLARE_REGRESSOR(GaussianRegressor, Gaussian);
GPR_DECLARE_REGRESSOR(LaplaceRegressor, Laplace);
GPR_DECLARE_REGRESSOR(ThinPlateRegressor, ThinPlate);

#undef GPR_DECLARE_REGRESSOR

}  // namespace gp_regression
