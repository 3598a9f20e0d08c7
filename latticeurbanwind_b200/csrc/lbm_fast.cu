// Contraction allowed (-fmad=true): what the reference's OpenCL build flags license (-cl-mad-enable, FX/opencl.hpp:305).
#define LUW_KERNELSET_FN kernels_fast
#define LUW_FAST true
#include "lbm_launch.inc"
