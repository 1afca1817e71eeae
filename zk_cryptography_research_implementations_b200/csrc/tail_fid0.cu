// tail kernels, field 0 (see tail_launch.cuh)
#define ZK_INSTANTIATE_TAIL 0
#include "tail_launch.cuh"
