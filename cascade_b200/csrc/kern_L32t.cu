// kernels for prec = 1024 bits: evaluation of solutions, exp / log, radius of convergence / truncation order
#include "cb_kernels.cuh"
CB_DEFINE_LAUNCH_TRANS(32)
