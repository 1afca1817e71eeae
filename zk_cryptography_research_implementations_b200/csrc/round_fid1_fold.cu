// fused fold + evaluation kernels, field 1 (see round_launch.cuh)
#define ZK_INSTANTIATE_FOLD_EVALS 1
#include "round_launch.cuh"
