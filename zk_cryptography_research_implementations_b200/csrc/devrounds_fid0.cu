// explicit instantiation of the device-resident round loop for field 0 (see devrounds_launch.cuh)
#define ZK_INSTANTIATE_DEVROUNDS 0
#include "devrounds_launch.cuh"
