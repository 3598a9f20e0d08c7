// Contraction allowed (-fmad=true): what the reference's OpenCL build flags license (-cl-mad-enable, FX/opencl.hpp:305).
#define LUW_KERNELSET_FN kernels_fast
#define LUW_FAST true
#include "lbm_launch.inc"

#ifdef LUW_TRACE // development aid (not part of the shipped library): copy the tile timestamps of the traced CTA to the host
extern "C" int luw_debug_trace(long long* host) { return (int)cudaMemcpyFromSymbol(host, luw::g_trace, sizeof(long long)*6*2048); }
#endif
