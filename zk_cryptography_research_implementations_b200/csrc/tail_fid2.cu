// tail kernels, field 2 (see tail_launch.cuh)
#define ZK_INSTANTIATE_TAIL 2
#include "tail_launch.cuh"
