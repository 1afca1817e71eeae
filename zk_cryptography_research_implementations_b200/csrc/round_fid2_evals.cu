// round-0 evaluation kernels, field 2 (see round_launch.cuh)
#define ZK_INSTANTIATE_ROUND_EVALS 2
#include "round_launch.cuh"
