// Scene-independent kernels (scheduler, state access, seeding, record packing) for both arithmetic types.
#include "lcr_kernels.cuh"
template struct lcr::Launch<float>;
template struct lcr::Launch<double>;
