// kernels for prec = 1024 bits: the Fuchs recurrences (table, shared / per-instance operator, integer operator)
#include "cb_kernels.cuh"
CB_DEFINE_LAUNCH_FUCHS(32)
