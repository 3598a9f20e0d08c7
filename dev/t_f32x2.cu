#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float2 a){ unsigned long long r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a.x),"f"(a.y)); return r;}
__device__ __forceinline__ float2 upk(unsigned long long a){ float2 r; asm("mov.b64 {%0,%1}, %2;":"=f"(r.x),"=f"(r.y):"l"(a)); return r;}
__global__ void k(float2* p) {
	float2 a = p[0], b = p[1], c = p[2];
	unsigned long long A=pk(a),B=pk(b),C=pk(c),R;
	asm("sub.rn.f32x2 %0, %1, %2;":"=l"(R):"l"(A),"l"(B));
	p[3] = upk(R);
	asm("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(R):"l"(A),"l"(B),"l"(C));
	p[4] = upk(R);
	float2 n = make_float2(-b.x,-b.y);
	p[5] = __fadd2_rn(a, n);
	p[6] = __ffma2_rn(a, n, c);
	p[7] = __fmul2_rn(a, make_float2(3.0f,3.0f));
	p[8] = __ffma2_rn(a, make_float2(c.x,c.x), b);
}
