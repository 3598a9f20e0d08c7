// TEST INFRASTRUCTURE (see cuda_on_host.hpp): the two inflow-sample search kernels of latticeurbanwind_b200/csrc/lbm_inlet.cuh, compiled for the host, behind the
// C-ABI names of include/luw_cuda.h. Linked INTO a test executable (tests/test_inlet_surface_on_host.py) these definitions take the place of libluw_cuda.so's, so that
// the host code around the kernels (latticeurbanwind_b200/host/inlet_outlet_surface.cpp) can be checked against the reference in the GPU-less container.
// Build: g++ -std=c++17 -O2 -ffp-contract=off -c -w -I/usr/local/cuda/include inlet_on_host.cpp
#include "cuda_on_host.hpp"
#include "../../latticeurbanwind_b200/csrc/lbm_inlet.cuh"
#include <vector>

static uint64_t emu_inlet_launches = 0ull;
template<class K> static void for_threads(const uint64_t n, const unsigned block, K kernel) {
	const unsigned gx = (unsigned)((n+block-1ull)/block);
	emu_blockDim = {block, 1u, 1u}; emu_gridDim = {gx, 1u, 1u};
	for(unsigned b=0u; b<gx; b++) for(unsigned t=0u; t<block; t++) { emu_blockIdx = {b, 0u, 0u}; emu_threadIdx = {t, 0u, 0u}; kernel(); }
	emu_inlet_launches++;
}
extern "C" {
int luw_inlet_nearest(int, uint64_t ncells, const float* cell_xyz, uint32_t npts, const float* point_xyz, uint32_t* nearest) {
	for_threads(ncells, 128u, [&]{ luw::k_inlet_nearest((uint32_t)ncells, cell_xyz, npts, point_xyz, nearest); });
	return 0;
}
int luw_inlet_knn(int, uint64_t ncells, const float* cell_ab, uint32_t npts, const float* point_ab, uint32_t* kept, uint32_t* used, float* max_r2, int32_t* exact) {
	for_threads(ncells, (unsigned)luw::INLET_KNN_THREADS, [&]{ luw::k_inlet_knn((uint32_t)ncells, cell_ab, npts, (const float2*)point_ab, kept, used, max_r2, exact); });
	return 0;
}
int luw_inlet_launch_count(uint64_t* launches) { *launches = emu_inlet_launches; return 0; }
}
