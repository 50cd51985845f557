// One translation unit per (arithmetic type, scene class): every kernel that depends on the scene class, fast and BIG
// workspace.  Compiled with -DLCR_T=float|double -DLCR_S=1|2|5 (see the Makefile).
#include "lcr_kernels.cuh"
template struct lcr::LaunchNC<LCR_T, LCR_S>;
