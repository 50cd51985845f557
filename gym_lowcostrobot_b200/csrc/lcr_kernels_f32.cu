// product precision: float32 state and arithmetic
#include "lcr_kernels.cuh"
template struct lcr::Launch<float>;
