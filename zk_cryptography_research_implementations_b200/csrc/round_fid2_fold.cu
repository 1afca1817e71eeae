// fused fold + evaluation kernels, field 2 (see round_launch.cuh)
#define ZK_INSTANTIATE_FOLD_EVALS 2
#include "round_launch.cuh"
