// Development micro-benchmark: what box size does a one-thread-issued TMA load/store pipeline need to saturate HBM on B200?
// Copies a 4-D fp16 tensor (x,y,z,slot) to another one, tile by tile: per tile 19 box loads into a shared-memory stage (one mbarrier),
// then 19 box stores from it. Loader and storer are two single threads in different warps; STAGES-deep ring; persistent CTAs.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mwait(uint64_t* b, uint32_t par) {
	uint32_t done = 0;
	while(!done) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 1000000;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(s32(b)), "r"(par) : "memory");
}
__global__ void __launch_bounds__(64, 1) copyk(const __grid_constant__ CUtensorMap src, const __grid_constant__ CUtensorMap dst, int tx, int ty, int tz, int bx, int by, int bz, int nslots, int slots_per_op, uint32_t box_bytes, int stages) {
	extern __shared__ __align__(128) uint8_t sm[];
	uint8_t* base = sm+((128u-(s32(sm)&127u))&127u);
	const uint32_t stage_bytes = box_bytes*nslots/slots_per_op*slots_per_op;
	uint64_t* full = (uint64_t*)(base+(size_t)stages*stage_bytes);
	uint64_t* empty = full+stages;
	if(threadIdx.x==0) { for(int s=0;s<stages;s++) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(full+s))); asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(empty+s))); } asm volatile("fence.mbarrier_init.release.cluster;"); }
	__syncthreads();
	const uint32_t ntiles = (uint32_t)tx*ty*tz;
	const int nops = nslots/slots_per_op;
	if(threadIdx.x==0) {
		uint32_t it = 0;
		for(uint32_t t=blockIdx.x; t<ntiles; t+=gridDim.x, it++) {
			const int s = it%stages;
			if(it>=(uint32_t)stages) mwait(empty+s, ((it/stages)-1)&1);
			const int x0 = (t%tx)*bx, y0 = ((t/tx)%ty)*by, z0 = (t/(tx*ty))*bz;
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(full+s)), "r"(stage_bytes) : "memory");
			for(int o=0; o<nops; o++) asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" :: "r"(s32(base+(size_t)s*stage_bytes+(size_t)o*box_bytes*slots_per_op)), "l"(&src), "r"(s32(full+s)), "r"(x0), "r"(y0), "r"(z0), "r"(o*slots_per_op) : "memory");
		}
	} else if(threadIdx.x==32) {
		uint32_t it = 0;
		for(uint32_t t=blockIdx.x; t<ntiles; t+=gridDim.x, it++) {
			const int s = it%stages;
			mwait(full+s, (it/stages)&1);
			const int x0 = (t%tx)*bx, y0 = ((t/tx)%ty)*by, z0 = (t/(tx*ty))*bz;
			for(int o=0; o<nops; o++) asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" :: "l"(&dst), "r"(s32(base+(size_t)s*stage_bytes+(size_t)o*box_bytes*slots_per_op)), "r"(x0), "r"(y0), "r"(z0), "r"(o*slots_per_op) : "memory");
			asm volatile("cp.async.bulk.commit_group;\ncp.async.bulk.wait_group.read 0;" ::: "memory");
			asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(s32(empty+s)) : "memory");
		}
		asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
	}
}
int main() {
	void* p = nullptr; cudaDriverEntryPointQueryResult q;
	cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
	enc_fn enc = (enc_fn)p;
	const uint64_t Nx=512, Ny=512, Nz=256, NS=18; // 18 slots so that 1,2,3,6,9 slots per op divide evenly
	const uint64_t N = Nx*Ny*Nz;
	uint16_t *a, *b; cudaMalloc(&a, NS*N*2); cudaMalloc(&b, NS*N*2); cudaMemset(a, 1, NS*N*2); cudaMemset(b, 0, NS*N*2);
	cudaFuncSetAttribute(copyk, cudaFuncAttributeMaxDynamicSharedMemorySize, 225*1024);
	struct C { int bx, by, bz, spo; } cfgs[] = { {64,4,2,1}, {64,8,2,1}, {64,8,4,1}, {128,8,4,1}, {64,4,2,2}, {64,4,2,3}, {64,4,2,9}, {64,8,2,9}, {128,4,1,1}, {256,2,1,1}, {256,4,1,1}, {256,8,1,1} };
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for(auto& c : cfgs) for(int ctas : {1, 2}) {
		CUtensorMap ms, md;
		cuuint64_t d4[4] = {Nx,Ny,Nz,NS}, s4[3] = {Nx*2, Nx*Ny*2, N*2}; cuuint32_t b4[4] = {(cuuint32_t)c.bx,(cuuint32_t)c.by,(cuuint32_t)c.bz,(cuuint32_t)c.spo}, e4[4] = {1,1,1,1};
		if(enc(&ms, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, a, d4, s4, b4, e4, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("encode failed\n"); continue; }
		enc(&md, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, b, d4, s4, b4, e4, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
		const uint32_t box_bytes = c.bx*c.by*c.bz*2;
		const uint32_t stage_bytes = box_bytes*NS;
		const int budget = (ctas==1 ? 220 : 108)*1024;
		int stages = budget/(int)stage_bytes; if(stages>12) stages = 12; if(stages<2) { printf("box %dx%dx%d x%d slots/op: stage too big\n", c.bx,c.by,c.bz,c.spo); continue; }
		const size_t smem = (size_t)stages*stage_bytes+stages*16+256;
		const int tx = Nx/c.bx, ty = Ny/c.by, tz = Nz/c.bz;
		float best = 1e9;
		for(int rep=0; rep<4; rep++) {
			cudaEventRecord(e0);
			copyk<<<148*ctas, 64, smem>>>(ms, md, tx, ty, tz, c.bx, c.by, c.bz, (int)NS, c.spo, box_bytes, stages);
			cudaEventRecord(e1); cudaEventSynchronize(e1);
			float ms_; cudaEventElapsedTime(&ms_, e0, e1); if(ms_<best) best = ms_;
		}
		cudaError_t e = cudaGetLastError();
		printf("box %3dx%dx%d (%5u B) x%d slot/op, %2d ops/tile, %2d stages, %d CTA/SM: %.3f ms  %.0f GB/s (read+write)  %s\n", c.bx, c.by, c.bz, box_bytes, c.spo, (int)NS/c.spo, stages, ctas, best, 2.0*NS*N*2/best/1e6, e==cudaSuccess ? "" : cudaGetErrorString(e));
	}
	return 0;
}
