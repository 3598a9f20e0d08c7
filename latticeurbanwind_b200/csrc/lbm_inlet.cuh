// Boundary-field evaluation on the device: the sample SEARCHES of the reference's inflow interpolators, one open-face cell per thread.
//
// The reference maps the wind samples (SurfData.csv: P points with a velocity each) onto the TYPE_E cells of the five open faces on host threads:
// NearestNeighborInterpolator::eval (FX/interpolation.cpp:53-62) scans all P samples per cell, KNNInterpolatorHD::eval (FX/interpolation_hd.cpp:184-411)
// scans them to keep the K = 64 nearest on the cell's face plane and then fits a weighted quadratic to those 64 in double precision. The scans are
// O(cells x P) -- 2 M face cells x 10^4..10^5 samples for a C3 / C4 lattice: minutes on the host -- and are what these kernels take over. The O(cells x K) fit
// stays on the host (latticeurbanwind_b200/host/inlet_outlet_surface.cpp): its weights are exp() in double, and only the host's libm gives the reference's bits.
//
// Bit-exactness: the selections below are the reference's sequential algorithms, run by one thread per cell in the reference's sample order, with every float
// product and sum individually rounded (__fmul_rn / __fadd_rn: nvcc may not contract them; the reference is compiled for x86-64 without FMA). Ties therefore
// resolve the same way, and the order of the kept samples -- the order the host's double-precision sums run in -- is the reference's.
// All threads of a warp read the same sample at the same time (one broadcast L1 transaction per warp and sample).
#pragma once
#include <cstdint>
#include <cfloat>

namespace luw {

constexpr int INLET_KNN_K = 64; // `constexpr int K = 64`, FX/interpolation_hd.cpp:185

// nearest[c] = index of the first sample with the smallest squared distance to cell c, 0xFFFFFFFF if no sample compares below FLT_MAX (-> u = 0).
// cell: SoA x[ncells] y[ncells] z[ncells]; pts: x, y, z per sample
__global__ void __launch_bounds__(128) k_inlet_nearest(const uint32_t ncells, const float* __restrict__ cell, const uint32_t npts, const float* __restrict__ pts, uint32_t* __restrict__ nearest) {
	const uint32_t c = blockIdx.x*blockDim.x+threadIdx.x;
	if(c>=ncells) return;
	const float px = cell[c], py = cell[(uint64_t)ncells+c], pz = cell[2ull*ncells+c];
	float best = FLT_MAX;
	uint32_t arg = 0xFFFFFFFFu;
	for(uint32_t i=0u; i<npts; i++) {
		const float dx = __fsub_rn(px, __ldg(pts+3ull*i)), dy = __fsub_rn(py, __ldg(pts+3ull*i+1ull)), dz = __fsub_rn(pz, __ldg(pts+3ull*i+2ull));
		const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)); // dot(d, d), FX/utilities.hpp:1027-1029
		if(d2<best) { best = d2; arg = i; }
	}
	nearest[c] = arg;
}

// The K-nearest selection of KNNInterpolatorHD::eval for cells that share a face plane. q: the in-plane coordinates (a, b) of the samples ON that plane, in the
// reference's sample order; cell: SoA a[ncells] b[ncells]. Per cell: exact[c] = first sample with r2 <= 1e-16 (the reference returns that sample's velocity) or -1;
// otherwise used[c] = number of kept samples (<= 64), kept[64*c + k] = their indices into q in the reference's slot order, max_r2[c] = `max_r2_kept`.
__global__ void __launch_bounds__(128) k_inlet_knn(const uint32_t ncells, const float* __restrict__ cell, const uint32_t npts, const float2* __restrict__ q,
	uint32_t* __restrict__ kept, uint32_t* __restrict__ used, float* __restrict__ max_r2, int32_t* __restrict__ exact) {
	constexpr int K = INLET_KNN_K;
	const uint32_t c = blockIdx.x*blockDim.x+threadIdx.x;
	if(c>=ncells) return;
	const float ca = cell[c], cb = cell[(uint64_t)ncells+c];
	float best_r2[K];
	uint32_t best_i[K];
	int filled = 0, worst_k = -1, hit = -1;
	float max_r2_kept = 0.0f, worst_r2 = 0.0f;
	for(uint32_t i=0u; i<npts; i++) {
		const float2 p = __ldg(q+i);
		const float s1 = __fsub_rn(p.x, ca), s2 = __fsub_rn(p.y, cb);
		const float r2 = __fadd_rn(__fmul_rn(s1, s1), __fmul_rn(s2, s2));
		if(r2<=1.0E-16f) { hit = (int)i; break; } // `eps2`
		if(filled<K) {
			best_r2[filled] = r2; best_i[filled] = i;
			if(r2>max_r2_kept) max_r2_kept = r2;
			filled++;
		} else {
			if(worst_k<0) { // the reference finds the worst kept sample anew for every candidate; it only changes when a sample is replaced
				worst_k = 0; worst_r2 = best_r2[0];
				for(int k=1; k<K; k++) if(best_r2[k]>worst_r2) { worst_r2 = best_r2[k]; worst_k = k; }
			}
			if(r2<worst_r2) {
				best_r2[worst_k] = r2; best_i[worst_k] = i;
				worst_k = 0; worst_r2 = best_r2[0];
				for(int k=1; k<K; k++) if(best_r2[k]>worst_r2) { worst_r2 = best_r2[k]; worst_k = k; }
				max_r2_kept = worst_r2;
			}
		}
	}
	exact[c] = hit;
	used[c] = (uint32_t)filled;
	max_r2[c] = max_r2_kept;
	for(int k=0; k<filled; k++) kept[(uint64_t)K*c+(uint64_t)k] = best_i[k];
}

} // namespace luw
