// kernels for prec = 2048 bits
#include "cb_kernels.cuh"
CB_DEFINE_LAUNCH_TABLE(64)
