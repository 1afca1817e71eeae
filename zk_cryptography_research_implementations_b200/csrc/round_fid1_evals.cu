// round-0 evaluation kernels, field 1 (see round_launch.cuh)
#define ZK_INSTANTIATE_ROUND_EVALS 1
#include "round_launch.cuh"
