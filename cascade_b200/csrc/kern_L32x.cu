// kernels for prec = 1024 bits: logarithmic Frobenius, solution bookkeeping, apply, operator and Taylor shifts
#include "cb_kernels.cuh"
CB_DEFINE_LAUNCH_OTHER(32)
