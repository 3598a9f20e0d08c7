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
constexpr int INLET_KNN_THREADS = 64; // threads per block of k_inlet_knn: 64 slots x (r2, index) + 8 group maxima, x 64 threads = 34 KB of shared memory
__global__ void __launch_bounds__(INLET_KNN_THREADS) k_inlet_knn(const uint32_t ncells, const float* __restrict__ cell, const uint32_t npts, const float2* __restrict__ q,
	uint32_t* __restrict__ kept, uint32_t* __restrict__ used, float* __restrict__ max_r2, int32_t* __restrict__ exact) {
	constexpr int K = INLET_KNN_K, T = INLET_KNN_THREADS, G = 8; // G groups of K/G slots
	// the kept samples of every thread, slot-major: slot k of thread t at [k*T + t] (consecutive threads, consecutive banks). In registers the slots could not be
	// indexed, in local memory every re-scan for the worst sample went through L1 / L2 to DRAM (first cut of this kernel: 31 GB of DRAM traffic per million cells).
	__shared__ float s_r2[K*T];
	__shared__ uint32_t s_i[K*T];
	__shared__ float s_gmax[G*T]; // largest r2 of each group of 8 slots (-inf if it holds NaN only): the worst sample is found in 8 + 8 loads instead of 64
	const uint32_t c = blockIdx.x*blockDim.x+threadIdx.x;
	if(c>=ncells) return;
	float* best_r2 = s_r2+threadIdx.x;
	uint32_t* best_i = s_i+threadIdx.x;
	float* gmax = s_gmax+threadIdx.x;
	const float ninf = __uint_as_float(0xFF800000u);
	// The reference's scan `worst = best[0]; for k = 1..63: if(best[k] > worst) ...` picks the FIRST slot that holds the largest value (a NaN in slot 0 stays the worst
	// for ever, NaN elsewhere is never picked). Same pick, in two levels: first group whose maximum is the largest, first slot of that group that holds it.
	const auto group_max = [&](const int g) { float m = ninf; for(int k=g*(K/G); k<(g+1)*(K/G); k++) { const float v = best_r2[k*T]; if(v>m) m = v; } gmax[g*T] = m; };
	const auto find_worst = [&](int& worst_k, float& worst_r2) {
		const float first = best_r2[0];
		if(first!=first) { worst_k = 0; worst_r2 = first; return; }
		int g = 0; float m = gmax[0];
		for(int j=1; j<G; j++) { const float v = gmax[j*T]; if(v>m) { m = v; g = j; } }
		worst_k = g*(K/G); worst_r2 = m; // m is a value some slot of group g holds (slot 0 is no NaN here, so group 0's maximum is no -inf and m neither)
		for(int k=g*(K/G); k<(g+1)*(K/G); k++) if(best_r2[k*T]==m) { worst_k = k; break; }
	};
	const float ca = cell[c], cb = cell[(uint64_t)ncells+c];
	int filled = 0, worst_k = -1, hit = -1;
	float max_r2_kept = 0.0f, worst_r2 = 0.0f;
	const auto dist2 = [&](const float a, const float b) { const float s1 = __fsub_rn(a, ca), s2 = __fsub_rn(b, cb); return __fadd_rn(__fmul_rn(s1, s1), __fmul_rn(s2, s2)); };
	// one sample of the reference's loop body; true: the sample coincides with the cell (`r2 <= eps2`), the scan ends
	const auto consider = [&](const uint32_t i, const float r2) {
		if(r2<=1.0E-16f) { hit = (int)i; return true; }
		if(filled<K) {
			best_r2[filled*T] = r2; best_i[filled*T] = i;
			if(r2>max_r2_kept) max_r2_kept = r2;
			filled++;
		} else {
			if(worst_k<0) { // the reference finds the worst kept sample anew for every candidate; it only changes when a sample is replaced
				for(int g=0; g<G; g++) group_max(g);
				find_worst(worst_k, worst_r2);
			}
			if(r2<worst_r2) {
				best_r2[worst_k*T] = r2; best_i[worst_k*T] = i;
				group_max(worst_k/(K/G));
				find_worst(worst_k, worst_r2);
				max_r2_kept = worst_r2;
			}
		}
		return false;
	};
	uint32_t i = 0u;
	// four samples per trip: their distances are independent of each other (the loop is latency-bound otherwise: three resident warps per scheduler); once the slots are
	// full almost every trip ends at the first test -- none of the four is nearer than the worst kept sample. Candidates go through the reference's body in sample order.
	for(; i+4u<=npts&&hit<0; i+=4u) {
		const float4 pa = __ldg((const float4*)(q+i)), pb = __ldg((const float4*)(q+i+2u));
		const float r0 = dist2(pa.x, pa.y), r1 = dist2(pa.z, pa.w), r2 = dist2(pb.x, pb.y), r3 = dist2(pb.z, pb.w);
		if(worst_k>=0&&!(r0<worst_r2||r1<worst_r2||r2<worst_r2||r3<worst_r2)&&r0>1.0E-16f&&r1>1.0E-16f&&r2>1.0E-16f&&r3>1.0E-16f) continue;
		if(consider(i, r0)||consider(i+1u, r1)||consider(i+2u, r2)||consider(i+3u, r3)) break;
	}
	for(; i<npts&&hit<0; i++) { const float2 p = __ldg(q+i); if(consider(i, dist2(p.x, p.y))) break; }
	exact[c] = hit;
	used[c] = (uint32_t)filled;
	max_r2[c] = max_r2_kept;
	for(int k=0; k<filled; k++) kept[(uint64_t)K*c+(uint64_t)k] = best_i[k*T];
}

} // namespace luw
