// Development probe: which cp.async.bulk.tensor forms does this B200 accept? (element-granular box origins, u8 3-D boxes, 4-D boxes)
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template<int RANK> __global__ void probe(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, int c3, uint32_t bytes, uint8_t* out, int do_store, int s0) {
	extern __shared__ __align__(128) uint8_t sm[];
	uint64_t* bar = (uint64_t*)(sm+65536);
	if(threadIdx.x==0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
	__syncthreads();
	if(threadIdx.x==0) {
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(bar)), "r"(bytes) : "memory");
		if(RANK==4) asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" :: "r"(s32(sm)), "l"(&map), "r"(s32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
		else asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" :: "r"(s32(sm)), "l"(&map), "r"(s32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
	}
	uint32_t done = 0, spins = 0;
	while(!done) { asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(s32(bar)) : "memory"); if(!done&&++spins>(1u<<22)) { if(threadIdx.x==0) printf("  timeout waiting for TMA\n"); return; } }
	for(uint32_t i=threadIdx.x; i<bytes; i+=blockDim.x) out[i] = sm[i];
	if(do_store) {
		for(uint32_t i=threadIdx.x; i<bytes; i+=blockDim.x) sm[i] = (uint8_t)(sm[i]+1);
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		__syncthreads();
		if(threadIdx.x==0) {
			if(RANK==4) asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" :: "l"(&map), "r"(s32(sm)), "r"(s0), "r"(c1), "r"(c2), "r"(c3) : "memory");
			else asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" :: "l"(&map), "r"(s32(sm)), "r"(s0), "r"(c1), "r"(c2) : "memory");
			asm volatile("cp.async.bulk.commit_group;\ncp.async.bulk.wait_group 0;" ::: "memory");
		}
	}
}
int main() {
	void* p = nullptr; cudaDriverEntryPointQueryResult q;
	cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
	enc_fn enc = (enc_fn)p;
	const uint32_t Nx=64, Ny=8, Nz=4;
	const uint64_t N = Nx*Ny*Nz;
	std::vector<float> h(19*N); for(size_t i=0;i<h.size();i++) h[i] = (float)i;
	float* d; cudaMalloc(&d, h.size()*4); cudaMemcpy(d, h.data(), h.size()*4, cudaMemcpyHostToDevice);
	std::vector<uint8_t> hf(N); for(size_t i=0;i<N;i++) hf[i] = (uint8_t)i;
	uint8_t* df; cudaMalloc(&df, N); cudaMemcpy(df, hf.data(), N, cudaMemcpyHostToDevice);
	uint8_t* out; cudaMalloc(&out, 65536);
	CUtensorMap m4, m3;
	cuuint64_t d4[4] = {Nx,Ny,Nz,19}, s4[3] = {Nx*4ull, Nx*Ny*4ull, N*4ull}; cuuint32_t b4[4] = {64,4,1,1}, e4[4] = {1,1,1,1};
	CUresult r = enc(&m4, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, d4, s4, b4, e4, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	printf("encode 4d f32: %d\n", (int)r);
	cuuint64_t d3[3] = {Nx,Ny,Nz}, s3[2] = {Nx, (cuuint64_t)Nx*Ny};
	r = enc(&m3, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, df, d3, s3, b4, e4, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	printf("encode 3d u8: %d\n", (int)r);
	cudaFuncSetAttribute(probe<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536+64);
	cudaFuncSetAttribute(probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536+64);
	std::vector<float> o(256);
	struct T { const char* name; int rank, c0, c1, c2, c3, store; } tests[] = {
		{"4d aligned (0,0,0,0)", 4, 0,0,0,0, 0}, {"4d slot 2 (0,0,0,2)", 4, 0,0,0,2, 0}, {"4d y-1 (0,-1,0,3)", 4, 0,-1,0,3, 0},
		{"4d x+4 (4,4,1,1)", 4, 4,4,1,1, 0}, {"3d u8 (0,0,0)", 3, 0,0,0,0, 0}, {"3d u8 (0,4,1)", 3, 0,4,1,0, 0}, {"4d aligned load+store", 4, 0,4,2,6, 1} };
	for(auto& t : tests) {
		cudaMemset(out, 0xEE, 65536);
		if(t.rank==4) probe<4><<<1,128,65536+64>>>(m4, t.c0,t.c1,t.c2,t.c3, 1024, out, t.store, t.c0); else probe<3><<<1,128,65536+64>>>(m3, t.c0,t.c1,t.c2,0, 256, out, t.store, t.c0);
		cudaError_t e = cudaDeviceSynchronize();
		printf("%-26s -> %s", t.name, cudaGetErrorString(e));
		if(e==cudaSuccess) {
			cudaMemcpy(o.data(), out, 1024, cudaMemcpyDeviceToHost);
			if(t.rank==4) printf("   first=%g  [63]=%g  [64]=%g (expect %g)", o[0], o[63], o[64], (float)(t.c3*N+(t.c2*Ny+t.c1)*Nx+t.c0));
			else printf("   first=%u", (unsigned)((uint8_t*)o.data())[0]);
		} else { printf("\n"); return 1; }
		printf("\n");
	}
	cudaMemcpy(h.data(), d, h.size()*4, cudaMemcpyDeviceToHost);
	printf("after stores: fi[5*N+(2*Ny+4)*Nx+1] = %g\n", h[5*N+(2*Ny+4)*Nx+1]);
	return 0;
}
