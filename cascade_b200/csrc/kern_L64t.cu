// kernels for prec = 2048 bits: evaluation of solutions, exp / log, radius of convergence / truncation order
#include "cb_kernels.cuh"
CB_DEFINE_LAUNCH_TRANS(64)
