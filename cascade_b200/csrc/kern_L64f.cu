// kernels for prec = 2048 bits: the Fuchs recurrences (table, shared / per-instance operator, integer operator)
#include "cb_kernels.cuh"
CB_DEFINE_LAUNCH_FUCHS(64)
