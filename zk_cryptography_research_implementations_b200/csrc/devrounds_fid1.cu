// explicit instantiation of the device-resident round loop for field 1 (see devrounds_launch.cuh)
#define ZK_INSTANTIATE_DEVROUNDS 1
#include "devrounds_launch.cuh"
