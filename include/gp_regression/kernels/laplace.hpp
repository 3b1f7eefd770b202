// Kept so that #include <gp_regression/kernels/laplace.hpp> keeps working; the class lives in cov_functions.h.
#pragma once
#include "../cov_functions.h"
