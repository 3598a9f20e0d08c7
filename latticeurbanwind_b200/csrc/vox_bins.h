// Host side of the binned voxeliser (k_voxelize_mesh_binned, lbm_kernels.cuh): which triangles can cross the rays of which 32 x 4 bin of columns. Plain C++ (no CUDA),
// shared by the C ABI (luw_cabi.cu) and the host-emulation tests.
//
// A ray along `direction` starts at the integer lattice coordinate (c0 + O0, c1 + O1) of its column (ray origin of FX/kernel.cpp:2391-2393: cell position + domain
// offset). It can cross a triangle only inside the triangle's projected bounding box; the box is padded by one cell on every side -- the margin the reference itself
// uses when it hands a domain only the triangles that can reach it (FX/lbm.cpp:41-90) -- which is four to five orders of magnitude above the rounding of the
// single-precision barycentric test at lattice coordinates FOR A WELL-CONDITIONED TRIANGLE. Triangles are entered in ascending order, so a bin's list is an ordered
// subsequence of the mesh.
//
// Conditioning. The kernel divides by g = u x v projected along the ray (twice the projected area). With g tiny against |u||v| -- a vertical sliver whose three corners
// project onto one line, up to rounding -- the rounding noise of the numerators, divided by g, reaches O(1), and the reference's test then "hits" for rays that lie near
// the sliver's line but several sliver lengths away from it. Those hits are noise, but they are the reference's flags. Such triangles therefore go into EVERY bin (what
// the all-triangles kernel does for all of them). Bound: a ray outside the padded box misses a barycentric constraint by m >= pad / edge; the computed coordinate is off
// by about 3 eps |w| / (kappa edge) with kappa = |g| / (|u||v|) and |w| <= the lattice diagonal (~1.5e4 cells): kappa >= 0.05 keeps the error below m / 20; the third
// constraint needs the projected edge b - c not to be short against |u| + |v|, tested the same way. Triangles with g == 0 EXACTLY (the walls of an extruded footprint:
// both corners of a vertical edge share their projection, the two products cancel bit for bit) fail the kernel's `g != 0` and are in no bin at all.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace luw {

constexpr uint32_t VOX_BIN0 = 32u, VOX_BIN1 = 4u; // columns per bin along the fast / slow axis of the face = the thread layout of one block

struct VoxBins {
	uint32_t bins0 = 0u, bins1 = 0u; // bins along the two axes of the face (0: no bin grid was built, see below)
	std::vector<uint32_t> start; // bins0*bins1 + 1 offsets into ids
	std::vector<uint32_t> ids; // triangle numbers, ascending within a bin
};

// face axes of a ray direction: x-rays (y, z), y-rays (z, x), z-rays (x, y)
inline void vox_face_axes(const uint32_t direction, int& a0, int& a1) { a0 = direction==0u ? 1 : direction==1u ? 2 : 0; a1 = direction==0u ? 2 : direction==1u ? 0 : 1; }

inline VoxBins vox_build_bins(const uint32_t direction, const uint32_t Nx, const uint32_t Ny, const uint32_t Nz, const int Ox, const int Oy, const int Oz,
	const float* p0, const float* p1, const float* p2, const uint32_t ntri) {
	int a0, a1;
	vox_face_axes(direction, a0, a1);
	const uint32_t N[3] = { Nx, Ny, Nz };
	const int O[3] = { Ox, Oy, Oz };
	const uint32_t n0 = N[a0], n1 = N[a1];
	VoxBins b;
	b.bins0 = (n0+VOX_BIN0-1u)/VOX_BIN0; b.bins1 = (n1+VOX_BIN1-1u)/VOX_BIN1;
	const uint64_t nbins = (uint64_t)b.bins0*b.bins1;
	b.start.assign(nbins+1ull, 0u);
	struct Range { uint32_t lo0, hi0, lo1, hi1; }; // bin ranges, inclusive; lo0 > hi0: the triangle reaches no column of this lattice
	std::vector<Range> range(ntri);
	uint64_t entries = 0ull;
	// columns [lo, hi] of an axis that a triangle spanning [tmin, tmax] (lattice coordinates) can reach, in bins of `width`
	const auto reach = [](const float tmin, const float tmax, const int offset, const uint32_t n, const uint32_t width, uint32_t& lo, uint32_t& hi) {
		const double first = std::ceil((double)tmin-1.0-(double)offset), last = std::floor((double)tmax+1.0-(double)offset);
		if(!(first<=last)||last<0.0||first>(double)(n-1u)) { lo = 1u; hi = 0u; return; }
		lo = (uint32_t)(first<0.0 ? 0.0 : first)/width; hi = (uint32_t)(last>(double)(n-1u) ? (double)(n-1u) : last)/width;
	};
	for(uint32_t i=0u; i<ntri; i++) {
		const float* v[3] = { p0+3ull*i, p1+3ull*i, p2+3ull*i };
		bool finite = true;
		for(int k=0; k<3; k++) for(int j=0; j<3; j++) finite = finite&&std::isfinite(v[k][j]); // a triangle with a NaN / Inf corner fails every comparison of the ray test
		Range r = { 1u, 0u, 1u, 0u };
		if(finite) {
			// u = p1 - p0, v = p2 - p0 as the kernel forms them; g = u_a0 v_a1 - u_a1 v_a0 up to sign (FX/kernel.cpp:2409-2413 with a unit ray direction)
			const float u0 = v[1][a0]-v[0][a0], u1 = v[1][a1]-v[0][a1], w0 = v[2][a0]-v[0][a0], w1 = v[2][a1]-v[0][a1];
			const double pa = (double)u0*(double)w1, pb = (double)u1*(double)w0; // exact in double
			if((float)pa==(float)pb) { range[i] = r; continue; } // the kernel's g is the difference of these two rounded products: exactly zero, the triangle cannot be hit
			const double lu = std::sqrt((double)u0*u0+(double)u1*u1), lv = std::sqrt((double)w0*w0+(double)w1*w1), lbc = std::sqrt(((double)w0-u0)*((double)w0-u0)+((double)w1-u1)*((double)w1-u1));
			const bool ill = !(std::fabs(pa-pb)>=0.05*lu*lv)||!(lbc>=0.05*(lu+lv));
			if(ill) { // see "Conditioning" above: every column tests it
				r = Range{ 0u, b.bins0-1u, 0u, b.bins1-1u };
				range[i] = r;
				entries += nbins;
				if(entries>0x7FFFFFFFull) { b.bins0 = b.bins1 = 0u; b.start.clear(); return b; }
				for(uint64_t k=0ull; k<nbins; k++) b.start[k+1ull]++;
				continue;
			}
			const float min0 = std::fmin(v[0][a0], std::fmin(v[1][a0], v[2][a0])), max0 = std::fmax(v[0][a0], std::fmax(v[1][a0], v[2][a0]));
			const float min1 = std::fmin(v[0][a1], std::fmin(v[1][a1], v[2][a1])), max1 = std::fmax(v[0][a1], std::fmax(v[1][a1], v[2][a1]));
			reach(min0, max0, O[a0], n0, VOX_BIN0, r.lo0, r.hi0);
			reach(min1, max1, O[a1], n1, VOX_BIN1, r.lo1, r.hi1);
			if(r.lo1>r.hi1) { r.lo0 = 1u; r.hi0 = 0u; }
		}
		range[i] = r;
		if(r.lo0<=r.hi0) {
			entries += (uint64_t)(r.hi1-r.lo1+1u)*(r.hi0-r.lo0+1u);
			if(entries>0x7FFFFFFFull) { b.bins0 = b.bins1 = 0u; b.start.clear(); return b; } // lists beyond 32-bit offsets: the caller takes the unbinned kernel
			for(uint32_t j=r.lo1; j<=r.hi1; j++) for(uint32_t k=r.lo0; k<=r.hi0; k++) b.start[(uint64_t)j*b.bins0+k+1ull]++;
		}
	}
	for(uint64_t k=0ull; k<nbins; k++) b.start[k+1ull] += b.start[k];
	b.ids.resize(b.start[nbins]);
	std::vector<uint32_t> cursor(b.start.begin(), b.start.end()-1);
	for(uint32_t i=0u; i<ntri; i++) {
		const Range& r = range[i];
		if(r.lo0<=r.hi0) for(uint32_t j=r.lo1; j<=r.hi1; j++) for(uint32_t k=r.lo0; k<=r.hi0; k++) b.ids[cursor[(uint64_t)j*b.bins0+k]++] = i;
	}
	return b;
}

} // namespace luw
