// fused fold + evaluation kernels, field 0 (see round_launch.cuh)
#define ZK_INSTANTIATE_FOLD_EVALS 0
#include "round_launch.cuh"
