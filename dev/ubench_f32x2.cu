// Micro-benchmark (development aid): FP32 issue throughput on sm_100a, scalar FFMA vs packed FFMA2 (fma.rn.f32x2), with and without
// interleaved ALU-pipe work. Prints lane-FMAs per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 4096
template<int MODE> __global__ void k(float* out, float a, float b, long long* cyc) {
	float x[16];
	for(int i=0;i<16;i++) x[i] = threadIdx.x*0.001f+i;
	unsigned u = threadIdx.x;
	long long t0 = clock64();
	for(int it=0; it<ITER; it++) {
		if(MODE==0) { // 16 scalar FFMA
#pragma unroll
			for(int i=0;i<16;i++) x[i] = fmaf(x[i], a, b);
		} else if(MODE==1) { // 8 FFMA2
#pragma unroll
			for(int i=0;i<16;i+=2) { float2 r = __ffma2_rn(make_float2(x[i],x[i+1]), make_float2(a,a), make_float2(b,b)); x[i]=r.x; x[i+1]=r.y; }
		} else if(MODE==2) { // 8 FFMA2 + 8 LOP3
#pragma unroll
			for(int i=0;i<16;i+=2) { float2 r = __ffma2_rn(make_float2(x[i],x[i+1]), make_float2(a,a), make_float2(b,b)); x[i]=r.x; x[i+1]=r.y; u = (u^(u<<1))&0x7fffffffu|(unsigned)i; }
		} else if(MODE==3) { // 16 FFMA + 8 LOP3
#pragma unroll
			for(int i=0;i<16;i++) { x[i] = fmaf(x[i], a, b); if(i&1) u = (u^(u<<1))&0x7fffffffu|(unsigned)i; }
		} else if(MODE==4) { // 8 FADD2
#pragma unroll
			for(int i=0;i<16;i+=2) { float2 r = __fadd2_rn(make_float2(x[i],x[i+1]), make_float2(a,b)); x[i]=r.x; x[i+1]=r.y; }
		} else if(MODE==5) { // 16 FADD
#pragma unroll
			for(int i=0;i<16;i++) x[i] = x[i]+a;
		}
	}
	long long t1 = clock64();
	float s = 0; for(int i=0;i<16;i++) s += x[i];
	out[blockIdx.x*blockDim.x+threadIdx.x] = s+u;
	if(threadIdx.x==0&&blockIdx.x==0) *cyc = t1-t0;
}
template<int MODE> void run(const char* name, int warps) {
	float* out; long long* cyc; cudaMalloc(&out, 148*1024*4); cudaMalloc(&cyc, 8);
	k<MODE><<<148, warps*32>>>(out, 1.0001f, 0.5f, cyc);
	k<MODE><<<148, warps*32>>>(out, 1.0001f, 0.5f, cyc);
	cudaDeviceSynchronize();
	long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
	printf("%-22s warps/SM=%2d  cycles=%lld  lane-FMA/clk/SM=%.1f\n", name, warps, c, 16.0*ITER*warps*32/(double)c);
	cudaFree(out); cudaFree(cyc);
}
int main() {
	for(int w : {8, 16, 32}) {
		run<0>("16xFFMA", w); run<1>("8xFFMA2", w); run<2>("8xFFMA2+8xLOP3", w); run<3>("16xFFMA+8xLOP3", w); run<4>("8xFADD2", w); run<5>("16xFADD", w);
	}
	return 0;
}
