// kernels for prec = 4096 bits
#include "cb_kernels.cuh"
CB_DEFINE_LAUNCH_TABLE(128)
