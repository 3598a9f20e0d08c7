#include <cuda_runtime.h>
__global__ void k(float2* p) {
	float2 a = p[0], b = p[1], c = p[2];
	p[3] = __fadd2_rn(__ffma2_rn(a, b, make_float2(-0.0f,-0.0f)), c);   // product rounded by an fma with -0: can it still be fused with the add?
	float2 m = make_float2(__fmul_rn(a.x,b.x), __fmul_rn(a.y,b.y));
	p[4] = __fadd2_rn(m, c);
}
