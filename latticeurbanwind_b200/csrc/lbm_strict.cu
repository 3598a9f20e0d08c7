// "As written" arithmetic: compiled with -fmad=false so that products and sums are rounded separately and only the explicit
// fmaf() calls are fused; IEEE division and square root. Bit-identical to oracle/ (and thereby to the reference kernel text).
#define LUW_KERNELSET_FN kernels_strict
#include "lbm_launch.inc"
