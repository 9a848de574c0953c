// kernels for prec = 1024 bits
#include "cb_kernels.cuh"
CB_DEFINE_LAUNCH_TABLE(32)
