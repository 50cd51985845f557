// verification build: the same kernels in float64 (parity against the CPU oracle to ~1e-9)
#include "lcr_kernels.cuh"
template struct lcr::Launch<double>;
