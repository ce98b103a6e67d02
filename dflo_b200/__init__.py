"""dflo_b200: B200-native explicit DG residual / RK-stage engine behind dflo's ConservationLaw
driver.  The compute path is the CUDA library dflo_b200/csrc/libdflo_b200.so, reached only
through the C ABI of include/dflo_b200.h; this package is the ctypes plumbing around it."""
from .abi import (BC, FLUX, DfloError, Engine, Mesh, Params, expr_eval, load_library, make_params)  # noqa: F401
