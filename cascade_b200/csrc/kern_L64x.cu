// kernels for prec = 2048 bits: logarithmic Frobenius, solution bookkeeping, apply, operator and Taylor shifts
#include "cb_kernels.cuh"
CB_DEFINE_LAUNCH_OTHER(64)
