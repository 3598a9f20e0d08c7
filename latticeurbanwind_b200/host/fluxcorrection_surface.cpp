// apply_flux_correction (FX/fluxcorrection.cpp:29-194) in O(surface) host time.
//
// What the reference does: it marks every non-solid cell of the five open faces (top, west, east, south, north; z > 0) TYPE_E, optionally re-evaluates the inflow on the
// downstream face, sums the outward normal velocity over those cells and subtracts the mean from every one of them, so that the boundary field is divergence-free in the
// integral sense. It finds the face cells by visiting ALL Nx*Ny*Nz cells through the stitched accessors (parallel_for over the volume, FX/fluxcorrection.cpp:54-73) --
// minutes of host time at 10^8 .. 10^10 cells for O(N^(2/3)) cells of work.
//
// Here the face cells are enumerated directly, in the order the reference's list ends up in: its parallel_for hands every thread a contiguous, ascending range of n and the
// per-thread lists are concatenated by thread number (FX/utilities.hpp:81-87, FX/fluxcorrection.cpp:75-81), i.e. ascending n whatever the thread count. The sums are
// accumulated in double in that same order, so `delta`, every corrected velocity, the flags and the three reported sums are bit-identical to the reference's
// (baseline/flux_parity.cpp runs both on the same lattice; tests/test_reference_driver.py). Same signature: in the drop-in driver this file takes the place of
// FX/fluxcorrection.cpp (baseline/build_reference_driver.py); the unmodified one remains the oracle.
#include "fluxcorrection.hpp"
#include <cmath>
#include <thread>
#include <vector>

namespace {

enum Face : unsigned char { TOP = 0, WEST, EAST, SOUTH, NORTH, FACES };
struct FaceCell { ulong n; Face face; };

// outward normal velocity of a face cell: component and sign per face
inline float& normal_component(LBM& lbm, const ulong n, const Face f) { return f==TOP ? lbm.u.z[n] : (f==WEST||f==EAST) ? lbm.u.x[n] : lbm.u.y[n]; }
inline float outward_sign(const Face f) { return (f==WEST||f==SOUTH) ? -1.0f : 1.0f; }

// the face cells with z in [z0, z1), ascending n. Precedence of the reference's pick_face: top, then x faces, then y faces (an edge cell belongs to the x face).
template<class Visit> void for_face_cells(const uint Nx, const uint Ny, const uint Nz, const uint z0, const uint z1, Visit&& visit) {
	for(uint z=(z0<1u ? 1u : z0); z<z1; z++) { // z == 0 is never a boundary cell of this routine
		if(z==Nz-1u) {
			for(uint y=0u; y<Ny; y++) for(uint x=0u; x<Nx; x++) visit(x, y, z, TOP);
			continue;
		}
		for(uint y=0u; y<Ny; y++) {
			if(y==0u||y==Ny-1u) {
				for(uint x=0u; x<Nx; x++) visit(x, y, z, x==0u ? WEST : x==Nx-1u ? EAST : y==0u ? SOUTH : NORTH);
			} else {
				visit(0u, y, z, WEST);
				if(Nx>1u) visit(Nx-1u, y, z, EAST);
			}
		}
	}
}

} // namespace

void apply_flux_correction(LBM& lbm, const std::string& downstream_bc, const std::function<float3(const float3&)>& inlet_eval, bool show_report,
	double* avg_delta_mps, double* net_before, double* net_after) {
	const uint Nx = lbm.get_Nx(), Ny = lbm.get_Ny(), Nz = lbm.get_Nz();
	const bool refill = static_cast<bool>(inlet_eval);
	const int down = downstream_bc=="+y" ? NORTH : downstream_bc=="-y" ? SOUTH : downstream_bc=="+x" ? EAST : downstream_bc=="-x" ? WEST : -1;
	const auto on_downstream_plane = [&](const uint x, const uint y) { return down==NORTH ? y==Ny-1u : down==SOUTH ? y==0u : down==EAST ? x==Nx-1u : down==WEST ? x==0u : false; };

	// 1. mark + (optionally) refill, z slabs in parallel; the slab lists, concatenated in slab order, are the reference's list
	const uint hw = std::thread::hardware_concurrency();
	const uint threads = (ulong)Nx*Ny*Nz<4096ull ? 1u : (hw==0u ? 1u : hw);
	(void)lbm.flags[0]; (void)lbm.u.x[0]; // the lazy host mirrors come into being on the calling thread
	std::vector<std::vector<FaceCell>> part(threads);
	{
		std::vector<std::thread> pool;
		for(uint t=0u; t<threads; t++) pool.emplace_back([&, t]() {
			const uint za = (uint)((ulong)Nz*t/threads), zb = (uint)((ulong)Nz*(t+1u)/threads);
			for_face_cells(Nx, Ny, Nz, za, zb, [&](const uint x, const uint y, const uint z, const Face f) {
				const ulong n = lbm.index(x, y, z);
				const uchar fl = lbm.flags[n];
				if((fl&TYPE_S)!=0u) return; // solids stay (terrain-clipped side cells included)
				lbm.flags[n] = (uchar)(fl|TYPE_E); // TYPE_T and auxiliary bits survive
				part[t].push_back(FaceCell{ n, f });
				if(refill&&on_downstream_plane(x, y)) {
					const float3 v = inlet_eval(lbm.position(x, y, z));
					lbm.u.x[n] = v.x; lbm.u.y[n] = v.y; lbm.u.z[n] = v.z;
				}
			});
		});
		for(std::thread& th : pool) th.join();
	}
	std::vector<FaceCell> cells;
	{ size_t total = 0u; for(const auto& p : part) total += p.size(); cells.reserve(total); }
	for(const auto& p : part) cells.insert(cells.end(), p.begin(), p.end());

	// 2. net outward flux, in list order, in double
	double inflow = 0.0, outflow = 0.0, net = 0.0;
	for(const FaceCell& c : cells) {
		const float vn = outward_sign(c.face)*normal_component(lbm, c.n, c.face);
		net += (double)vn;
		if(vn<0.0f) inflow += (double)(-vn); else outflow += (double)vn;
	}
	if(net_before) *net_before = net;
	const ulong count = (ulong)cells.size();
	const double delta = count>0ull ? -net/(double)count : 0.0;

	// 3. subtract the mean from every face cell's normal component
	double moved_sum = 0.0, moved_face[FACES] = { 0.0, 0.0, 0.0, 0.0, 0.0 };
	ulong moved = 0ull, moved_count[FACES] = { 0ull, 0ull, 0ull, 0ull, 0ull };
	for(const FaceCell& c : cells) {
		const float ox = lbm.u.x[c.n], oy = lbm.u.y[c.n], oz = lbm.u.z[c.n];
		float& comp = normal_component(lbm, c.n, c.face);
		comp = comp+outward_sign(c.face)*(float)delta;
		const float dx = lbm.u.x[c.n]-ox, dy = lbm.u.y[c.n]-oy, dz = lbm.u.z[c.n]-oz;
		const double d = std::sqrt((double)dx*dx+(double)dy*dy+(double)dz*dz);
		moved_sum += d; moved_face[c.face] += d; moved_count[c.face]++; moved++;
	}

	// 4. what is left
	double after = 0.0;
	for(const FaceCell& c : cells) after += (double)(outward_sign(c.face)*normal_component(lbm, c.n, c.face));
	if(avg_delta_mps) *avg_delta_mps = moved>0ull ? moved_sum/(double)moved : 0.0;
	if(net_after) *net_after = after;

	if(show_report) {
		const auto mean = [&](const Face f) { return moved_count[f] ? moved_face[f]/(double)moved_count[f] : 0.0; };
		println("| Flux correction | S_in="+to_string(inflow, 3u)+", S_out="+to_string(outflow, 3u)+", net_before="+to_string(net, 3u)+" |");
		println("| Flux correction | avg_dU="+to_string(delta, 3u)+" m/s, corrected="+to_string(count)+", net_after="+to_string(after, 3u)+" |");
		println("| Flux correction | per-face dU: Xn="+to_string(mean(WEST), 3u)+", Xp="+to_string(mean(EAST), 3u)+", Yn="+to_string(mean(SOUTH), 3u)+", Yp="+to_string(mean(NORTH), 3u)+", Zp="+to_string(mean(TOP), 3u)+" m/s |");
	}
}
