// tail kernels, field 1 (see tail_launch.cuh)
#define ZK_INSTANTIATE_TAIL 1
#include "tail_launch.cuh"
