// explicit instantiation of the device-resident round loop for field 2 (see devrounds_launch.cuh)
#define ZK_INSTANTIATE_DEVROUNDS 2
#include "devrounds_launch.cuh"
