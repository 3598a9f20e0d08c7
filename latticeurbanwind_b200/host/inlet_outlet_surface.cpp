// apply_inlet_outlet (FX/interpolation.cpp:66-210) and apply_inlet_outlet_hd (FX/interpolation_hd.cpp:443-750): O(surface) cell enumeration on the
// host, the O(cells x samples) searches on the GPU.
//
// Both reference routines build the boundary of a case: the ground plane z = 0 becomes TYPE_S, every cell of the five open faces becomes TYPE_E and -- unless it lies on an
// open downstream face -- gets the inflow velocity of its position (nearest wind sample / K-nearest-neighbour quadratic fit; side cells above `side_ref_z_cap_index` are
// evaluated at that height). They find those cells by visiting ALL Nx*Ny*Nz cells through the stitched accessors, and evaluate each one by scanning ALL samples on host
// threads: minutes at 10^8 .. 10^10 cells. Every write depends on the cell alone (no sums across cells, no order), so
//   1. the ground plane and the faces are enumerated directly (O(N^(2/3)) instead of O(N) host work),
//   2. the sample scans run on the device, one face cell per thread, as the reference's own sequential loops (csrc/lbm_inlet.cuh: luw_inlet_nearest, luw_inlet_knn),
//   3. what is left per cell on the host is O(1) (nearest: copy the sample's velocity) or O(K = 64) (HD: the weighted quadratic fit, in the reference's operation order,
//      with the host's exp() -- the one function whose bits a device cannot reproduce),
// and the flags and velocities come out bit-identical to the reference's. Interpolators this file does not know (ConstantInletInterpolator, a user's own subclass) are
// evaluated through their virtual eval() on host threads, still surface-only.
//
// In the drop-in driver FX/interpolation.cpp and FX/interpolation_hd.cpp are compiled unmodified with their two apply_* functions renamed
// (-Dapply_inlet_outlet=ref_apply_inlet_outlet ...: the interpolator classes stay the reference's), and this file supplies the two names
// (baseline/build_reference_driver.py). baseline/inlet_parity.cpp runs the renamed originals and these on identical lattices and requires identical flags and u
// (tests/test_reference_driver.py, on the GPU); tests/test_inlet_surface_on_host.py checks the same code against the reference's eval() in the GPU-less container, with
// the kernel source compiled for the host.
#include "setup.hpp" // everything the two headers below include, before `private` is redefined
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <thread>
#include <vector>
// The interpolators keep their samples private and offer eval() only. The device search needs the sample arrays themselves; the class layouts are untouched.
#define private public
#include "interpolation.hpp"
#include "interpolation_hd.hpp"
#undef private

namespace {

struct SurfaceCell { ulong n; float3 pos; }; // lattice index + the position its velocity is evaluated at

unsigned worker_threads() { // the reference's rule: all hardware threads, LBM_NUM_THREADS may lower it (FX/interpolation.cpp:76-96)
	const unsigned hw0 = std::thread::hardware_concurrency(), hw = hw0==0u ? 4u : hw0;
	if(const char* e = std::getenv("LBM_NUM_THREADS")) { const long v = std::strtol(e, nullptr, 10); if(v>0) return (unsigned)(v<(long)hw ? v : (long)hw); }
	return hw;
}
template<class Work> void parallel_over(const size_t count, Work&& work) { // work(i) for i in [0, count), every i independent
	const unsigned threads = count<4096u ? 1u : worker_threads();
	std::atomic<size_t> next{ 0u };
	const size_t chunk = count/((size_t)threads*32u)+64u;
	const auto loop = [&]() {
		for(;;) {
			const size_t a = next.fetch_add(chunk, std::memory_order_relaxed);
			if(a>=count) break;
			const size_t b = a+chunk<count ? a+chunk : count;
			for(size_t i=a; i<b; i++) work(i);
		}
	};
	std::vector<std::thread> pool;
	for(unsigned t=1u; t<threads; t++) pool.emplace_back(loop);
	loop();
	for(std::thread& th : pool) th.join();
}

// ground plane -> TYPE_S; face cells (z > 0) -> TYPE_E; returns those that take an inflow velocity, with their evaluation position
std::vector<SurfaceCell> mark_boundary(LBM& lbm, const std::string& downstream_bc, const bool downstream_open_face, const int side_ref_z_cap_index) {
	const uint Nx = lbm.get_Nx(), Ny = lbm.get_Ny(), Nz = lbm.get_Nz();
	for(uint y=0u; y<Ny; y++) for(uint x=0u; x<Nx; x++) lbm.flags[lbm.index(x, y, 0u)] = TYPE_S;
	const int down = downstream_bc=="+y" ? 3 : downstream_bc=="-y" ? 2 : downstream_bc=="+x" ? 1 : downstream_bc=="-x" ? 0 : -1;
	const auto outlet = [&](const uint x, const uint y) { return down==3 ? y==Ny-1u : down==2 ? y==0u : down==1 ? x==Nx-1u : down==0 ? x==0u : false; };
	std::vector<SurfaceCell> fill;
	const auto visit = [&](const uint x, const uint y, const uint z) {
		const ulong n = lbm.index(x, y, z);
		lbm.flags[n] = TYPE_E;
		if(downstream_open_face&&outlet(x, y)) return;
		float3 pos = lbm.position(x, y, z);
		const bool side = x==0u||x==Nx-1u||y==0u||y==Ny-1u;
		if(side_ref_z_cap_index>=0&&side&&z!=Nz-1u&&(int)z>side_ref_z_cap_index) pos.z = lbm.position(x, y, (uint)side_ref_z_cap_index).z;
		fill.push_back(SurfaceCell{ n, pos });
	};
	for(uint z=1u; z<Nz; z++) {
		if(z==Nz-1u) { for(uint y=0u; y<Ny; y++) for(uint x=0u; x<Nx; x++) visit(x, y, z); continue; }
		for(uint y=0u; y<Ny; y++) {
			if(y==0u||y==Ny-1u) { for(uint x=0u; x<Nx; x++) visit(x, y, z); }
			else { visit(0u, y, z); if(Nx>1u) visit(Nx-1u, y, z); }
		}
	}
	return fill;
}
void store(LBM& lbm, const std::vector<SurfaceCell>& cells, const std::vector<float3>& u) {
	for(size_t i=0u; i<cells.size(); i++) { const ulong n = cells[i].n; lbm.u.x[n] = u[i].x; lbm.u.y[n] = u[i].y; lbm.u.z[n] = u[i].z; }
}
int device_of(LBM& lbm) { return lbm.lbm_domain[0]->get_device_ordinal(); }
void check(const int rc) { if(rc!=LUW_OK) print_error(std::string("CUDA layer: ")+luw_last_error_string()); }

// ---- the K = 64 fit of KNNInterpolatorHD::eval (FX/interpolation_hd.cpp:298-410) over the samples the device kept, in the reference's operation order ----
struct KeptSample { double q1, q2; float3 u; };

// 6 unknowns, 3 right-hand sides: Gaussian elimination, partial pivoting on the largest |a[i][k]| (first one wins), pivots below 1e-18 give up (FX/interpolation_hd.cpp:57-152)
bool solve6(double a[6][6], double r[3][6], double x[3][6]) {
	for(int k=0; k<6; k++) {
		int pivot = k;
		double largest = std::fabs(a[k][k]);
		for(int i=k+1; i<6; i++) { const double v = std::fabs(a[i][k]); if(v>largest) { largest = v; pivot = i; } }
		if(largest<1e-18) return false;
		if(pivot!=k) {
			for(int j=0; j<6; j++) std::swap(a[k][j], a[pivot][j]);
			for(int c=0; c<3; c++) std::swap(r[c][k], r[c][pivot]);
		}
		const double inv_diag = 1.0/a[k][k];
		for(int i=k+1; i<6; i++) {
			const double factor = a[i][k]*inv_diag;
			if(factor==0.0) continue;
			for(int j=k; j<6; j++) a[i][j] -= factor*a[k][j];
			for(int c=0; c<3; c++) r[c][i] -= factor*r[c][k];
		}
	}
	for(int c=0; c<3; c++) for(int i=0; i<6; i++) x[c][i] = 0.0;
	for(int i=5; i>=0; i--) {
		double s[3] = { r[0][i], r[1][i], r[2][i] };
		for(int j=i+1; j<6; j++) for(int c=0; c<3; c++) s[c] -= a[i][j]*x[c][j];
		if(std::fabs(a[i][i])<1e-18) return false;
		const double inv_diag = 1.0/a[i][i];
		for(int c=0; c<3; c++) x[c][i] = s[c]*inv_diag;
	}
	return true;
}
float3 fit_kept(const KeptSample* s, const int used, const float max_r2_kept) {
	if(used==0) return float3(0.0f, 0.0f, 0.0f);
	const double R2 = (double)std::max(max_r2_kept, 1e-12f), sigma2 = 0.25*R2;
	if(used>=6) {
		double A[6][6] = {}, b[3][6] = {}, x[3][6];
		for(int kk=0; kk<used; kk++) {
			const double q1 = s[kk].q1, q2 = s[kk].q2, r2d = q1*q1+q2*q2, w = std::exp(-r2d/(2.0*sigma2));
			const double phi[6] = { 1.0, q1, q2, q1*q1, q1*q2, q2*q2 };
			for(int i=0; i<6; i++) { const double wi = w*phi[i]; for(int j=0; j<6; j++) A[i][j] += wi*phi[j]; }
			for(int i=0; i<6; i++) { const double wphi = w*phi[i]; b[0][i] += wphi*(double)s[kk].u.x; b[1][i] += wphi*(double)s[kk].u.y; b[2][i] += wphi*(double)s[kk].u.z; }
		}
		if(solve6(A, b, x)) return float3((float)x[0][0], (float)x[1][0], (float)x[2][0]);
	}
	double wx = 0.0, wy = 0.0, wz = 0.0, wsum = 0.0; // fewer than 6 samples, or a singular system: Gaussian-weighted mean
	for(int k=0; k<used; k++) {
		const double q1 = s[k].q1, q2 = s[k].q2, r2d = q1*q1+q2*q2, w = std::exp(-r2d/(2.0*sigma2));
		wx += w*(double)s[k].u.x; wy += w*(double)s[k].u.y; wz += w*(double)s[k].u.z; wsum += w;
	}
	if(wsum<=0.0) return float3(0.0f, 0.0f, 0.0f);
	const double inv = 1.0/wsum;
	return float3((float)(wx*inv), (float)(wy*inv), (float)(wz*inv));
}

} // namespace

// InletVelocityField over a NearestNeighborInterpolator (FX/interpolation.cpp:53-64) at `count` positions: u[i] = inlet(pos[i]), the sample scans on `device`
void luw_inlet_eval_nearest(const int device, const NearestNeighborInterpolator& nn, const float z_threshold, const float3* pos, const size_t count, float3* u) {
	std::vector<size_t> todo; // cells at or above the threshold: `if (z_lb < z0_ + zoff_) return float3(0.0f)`
	for(size_t i=0u; i<count; i++) { if(pos[i].z<z_threshold) u[i] = float3(0.0f); else todo.push_back(i); }
	const size_t n = todo.size();
	std::vector<float> cell(3u*n);
	for(size_t k=0u; k<n; k++) { const float3& p = pos[todo[k]]; cell[k] = p.x; cell[n+k] = p.y; cell[2u*n+k] = p.z; }
	std::vector<uint> nearest(n);
	check(luw_inlet_nearest(device, (uint64_t)n, cell.data(), (uint32_t)nn.P.size(), (const float*)nn.P.data(), nearest.data()));
	for(size_t k=0u; k<n; k++) u[todo[k]] = nearest[k]==0xFFFFFFFFu ? float3(0) : nn.U[nearest[k]];
}

// InletVelocityFieldHD over a KNNInterpolatorHD (FX/interpolation_hd.cpp:184-421) at `count` positions
void luw_inlet_eval_knn(const int device, const KNNInterpolatorHD& knn, const float z_base, const float3* pos, const size_t count, float3* u) {
	const std::vector<float3>& P = knn.P_;
	const std::vector<float3>& U = knn.U_;
	const int Pn = (int)P.size();
	std::vector<size_t> todo[5]; // cells at or above the threshold, by the face plane of the sample cloud they are closest to
	if(Pn==0) { for(size_t i=0u; i<count; i++) u[i] = float3(0.0f, 0.0f, 0.0f); return; }
	// bounding box of the samples, plane tolerance, plane of a position: FX/interpolation_hd.cpp:202-249 (the reference repeats this for every cell)
	float xmin = P[0].x, xmax = P[0].x, ymin = P[0].y, ymax = P[0].y, zmin = P[0].z, zmax = P[0].z;
	for(int i=1; i<Pn; i++) {
		const float x = P[i].x, y = P[i].y, z = P[i].z;
		if(x<xmin) xmin = x; if(x>xmax) xmax = x; if(y<ymin) ymin = y; if(y>ymax) ymax = y; if(z<zmin) zmin = z; if(z>zmax) zmax = z;
	}
	const float ex = xmax-xmin, ey = ymax-ymin, ez = zmax-zmin;
	float max_extent = ex; if(ey>max_extent) max_extent = ey; if(ez>max_extent) max_extent = ez;
	const float plane_tol = 1e-5f*max_extent+1e-6f;
	for(size_t i=0u; i<count; i++) {
		if(pos[i].z<z_base) { u[i] = float3(0.0f, 0.0f, 0.0f); continue; }
		const float d[5] = { std::fabs(pos[i].x-xmin), std::fabs(pos[i].x-xmax), std::fabs(pos[i].y-ymin), std::fabs(pos[i].y-ymax), std::fabs(pos[i].z-zmax) };
		int plane = 0; float dmin = d[0];
		for(int f=1; f<5; f++) if(d[f]<dmin) { dmin = d[f]; plane = f; }
		todo[plane].push_back(i);
	}
	for(int plane=0; plane<5; plane++) {
		const std::vector<size_t>& cells = todo[plane];
		const size_t n = cells.size();
		if(n==0u) continue;
		// the samples on this plane, in sample order, as in-plane coordinates (surface_local_coords, FX/interpolation_hd.cpp:155-181)
		const auto a_of = [plane](const float3& p) { return plane<2 ? p.y : p.x; };
		const auto b_of = [plane](const float3& p) { return plane<4 ? p.z : p.y; };
		std::vector<int> on_plane; std::vector<float> q;
		for(int i=0; i<Pn; i++) {
			const float dist = plane==0 ? std::fabs(P[i].x-xmin) : plane==1 ? std::fabs(P[i].x-xmax) : plane==2 ? std::fabs(P[i].y-ymin) : plane==3 ? std::fabs(P[i].y-ymax) : std::fabs(P[i].z-zmax);
			if(dist<=plane_tol) { on_plane.push_back(i); q.push_back(a_of(P[i])); q.push_back(b_of(P[i])); }
		}
		// in batches, so that the table of kept samples (256 B per cell) stays bounded on the host as well (LUW_INLET_BATCH: test hook for the batching itself)
		static const size_t batch = []{ const char* e = std::getenv("LUW_INLET_BATCH"); const long v = e ? std::atol(e) : 0l; return v>0l ? (size_t)v : (size_t)1u<<20; }();
		std::vector<float> cell(2u*std::min(n, batch));
		std::vector<uint> kept((size_t)LUW_INLET_KNN_K*std::min(n, batch)), used(std::min(n, batch)); std::vector<float> max_r2(std::min(n, batch)); std::vector<int> exact(std::min(n, batch));
		for(size_t c0=0u; c0<n; c0+=batch) {
			const size_t m = std::min(batch, n-c0);
			for(size_t k=0u; k<m; k++) { cell[k] = a_of(pos[cells[c0+k]]); cell[m+k] = b_of(pos[cells[c0+k]]); }
			check(luw_inlet_knn(device, (uint64_t)m, cell.data(), (uint32_t)on_plane.size(), q.data(), kept.data(), used.data(), max_r2.data(), exact.data()));
			parallel_over(m, [&](const size_t k) {
				float3& out = u[cells[c0+k]];
				if(exact[k]>=0) { out = U[on_plane[exact[k]]]; return; }
				KeptSample s[LUW_INLET_KNN_K];
				const float ca = cell[k], cb = cell[m+k];
				for(uint j=0u; j<used[k]; j++) {
					const uint slot = kept[(size_t)LUW_INLET_KNN_K*k+j];
					const float s1 = q[2u*slot]-ca, s2 = q[2u*slot+1u]-cb;
					s[j] = KeptSample{ (double)s1, (double)s2, U[on_plane[slot]] };
				}
				out = fit_kept(s, (int)used[k], max_r2[k]);
			});
		}
	}
}

#ifdef LUW_INLET_AB_HOOK
// TEST HOOK of the drop-in driver build (baseline/build_reference_driver.py defines the macro): with LUW_INLET_AB=reference in the environment the two functions below hand
// over to the reference's own, unmodified ones -- linked into the same binary under other names -- so that tests/test_reference_driver.py can run one deck both ways and
// compare the boundary images byte for byte. Both ways are host-side case set-up; the LBM step is not involved.
void ref_apply_inlet_outlet(LBM& lbm, const std::string& downstream_bc, const InletVelocityField& inlet, bool downstream_open_face, unsigned long min_work_per_thread, bool show_progress, int side_ref_z_cap_index);
void ref_apply_inlet_outlet_hd(LBM& lbm, const std::string& downstream_bc, const InletVelocityFieldHD& inlet, bool downstream_open_face, unsigned long min_work_per_thread, bool show_progress, int side_ref_z_cap_index);
static bool ab_reference() { const char* e = std::getenv("LUW_INLET_AB"); return e&&std::string(e)=="reference"; }
#endif

void apply_inlet_outlet(LBM& lbm, const std::string& downstream_bc, const InletVelocityField& inlet, bool downstream_open_face, unsigned long min_work_per_thread,
	bool show_progress, int side_ref_z_cap_index) {
#ifdef LUW_INLET_AB_HOOK
	if(ab_reference()) { println("| inlet/outlet init: LUW_INLET_AB=reference -> the reference's own apply_inlet_outlet"); ref_apply_inlet_outlet(lbm, downstream_bc, inlet, downstream_open_face, min_work_per_thread, show_progress, side_ref_z_cap_index); return; }
#endif
	(void)min_work_per_thread;
	const std::vector<SurfaceCell> cells = mark_boundary(lbm, downstream_bc, downstream_open_face, side_ref_z_cap_index);
	std::vector<float3> pos(cells.size()), u(cells.size());
	for(size_t i=0u; i<cells.size(); i++) pos[i] = cells[i].pos;
	const NearestNeighborInterpolator* nn = dynamic_cast<const NearestNeighborInterpolator*>(&inlet.interp_);
	if(nn) luw_inlet_eval_nearest(device_of(lbm), *nn, inlet.z0_+inlet.zoff_, pos.data(), pos.size(), u.data());
	else parallel_over(cells.size(), [&](const size_t i) { u[i] = inlet(pos[i]); });
	store(lbm, cells, u);
	if(show_progress) {
		if(luw_progress_gui_mode()) luw_emit_progress("interface_interpolation", "Interface interpolation", "Low-order inlet/outlet mapping completed", (long long)lbm.get_N(), (long long)lbm.get_N(), false);
		else println("| inlet/outlet init: "+to_string((ulong)cells.size())+" boundary cells mapped ("+string(nn ? "sample search on the GPU" : "host")+")");
	}
}

void apply_inlet_outlet_hd(LBM& lbm, const std::string& downstream_bc, const InletVelocityFieldHD& inlet, bool downstream_open_face, unsigned long min_work_per_thread,
	bool show_progress, int side_ref_z_cap_index) {
#ifdef LUW_INLET_AB_HOOK
	if(ab_reference()) { println("| inlet/outlet init (HD): LUW_INLET_AB=reference -> the reference's own apply_inlet_outlet_hd"); ref_apply_inlet_outlet_hd(lbm, downstream_bc, inlet, downstream_open_face, min_work_per_thread, show_progress, side_ref_z_cap_index); return; }
#endif
	(void)min_work_per_thread;
	const std::vector<SurfaceCell> cells = mark_boundary(lbm, downstream_bc, downstream_open_face, side_ref_z_cap_index);
	std::vector<float3> pos(cells.size()), u(cells.size());
	for(size_t i=0u; i<cells.size(); i++) pos[i] = cells[i].pos;
	const KNNInterpolatorHD* knn = dynamic_cast<const KNNInterpolatorHD*>(&inlet.interp_);
	if(knn) luw_inlet_eval_knn(device_of(lbm), *knn, inlet.z_base_lbmu_, pos.data(), pos.size(), u.data());
	else parallel_over(cells.size(), [&](const size_t i) { u[i] = inlet(pos[i]); });
	store(lbm, cells, u);
	if(show_progress) {
		if(luw_progress_gui_mode()) luw_emit_progress("interface_interpolation", "Interface interpolation", "High-order inlet/outlet mapping completed", 1ll, 1ll, false);
		else println("| inlet/outlet init (HD): "+to_string((ulong)cells.size())+" boundary cells mapped ("+string(knn ? "K-nearest search on the GPU, fit on the host" : "host")+")");
	}
}
