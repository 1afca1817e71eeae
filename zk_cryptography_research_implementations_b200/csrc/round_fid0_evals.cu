// round-0 evaluation kernels, field 0 (see round_launch.cuh)
#define ZK_INSTANTIATE_ROUND_EVALS 0
#include "round_launch.cuh"
