"""CPU oracle for the GP-regression hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this package.  The product path (the CUDA library behind include/gpr_c_api.h and the C++
drop-in headers) never does, and fails loudly when its CUDA library is missing.

Three checkers live here:

* ``Oracle``      — ctypes binding of oracle/gpr_oracle.cpp, our C++ restatement of
                    /root/reference/include/gp_regression/gp_regressor.hpp (double or long double,
                    pivoted LDLT as the reference or plain LLT, difference- or expansion-form distance).
* ``Reference``   — ctypes binding of oracle/_ref/libgpr_ref.so: the reference's own header compiled
                    unmodified against oracle/eigen_shim (Eigen is absent from the image).  Present
                    only if oracle/_ref was built where /root/reference exists.
* ``blas_*``      — the same mathematics with numpy/scipy (OpenBLAS LAPACK on all host cores): the
                    "best-effort CPU" flavour of BASELINE.md §5, used as the timed CPU baseline.
* ``certify``     — extended-precision a-posteriori certificate of mean / variance from ANY approximate
                    solves (residuals in long double, K never stored): the independent check at the
                    headline sizes n = 16 384 / 65 536 where no CPU factorisation fits a test's budget.
"""
from .oracle import (Oracle, Reference, build, have_reference, blas_fit, blas_predict,
                     kernel_value, tangent_basis, residual, certify)
