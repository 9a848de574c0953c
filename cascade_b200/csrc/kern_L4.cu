// kernels for prec = 128 bits
#include "cb_kernels.cuh"
CB_DEFINE_LAUNCH_TABLE(4)
