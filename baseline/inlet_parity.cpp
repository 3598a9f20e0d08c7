// TEST HARNESS for SURVEY.md 8-f2 (latticeurbanwind_b200/host/inlet_outlet_surface.cpp against the reference's own, unmodified code):
//   level 1, positions: InletVelocityField / InletVelocityFieldHD -- the reference's functors over its NearestNeighborInterpolator / KNNInterpolatorHD, compiled from
//            FX/interpolation.cpp / interpolation_hd.cpp where they lie -- against luw_inlet_eval_nearest / luw_inlet_eval_knn (sample search on the device, fit on the host)
//            at the same positions: every velocity must be bit-identical. Sample clouds: regular grids (ties everywhere, samples that coincide with cells), jittered
//            clouds, planes with 0 / 3 / 40 samples, collinear samples (singular fit -> weighted mean), no samples at all; with and without a z threshold inside the box.
//   level 2, lattices (needs a GPU: LBM objects): ref_apply_inlet_outlet(_hd) -- the reference's functions under another name -- against apply_inlet_outlet(_hd) of this
//            repo on two LBM objects: flags and u of ALL cells must be bit-identical, for every downstream face, open / closed, with / without the side z cap.
// baseline/build_reference_driver.py links it into baseline/_ref/luw_inlet_parity (tests/test_reference_driver.py runs it on the GPU box). Built with -DLUW_INLET_ON_HOST
// (tests/test_inlet_surface_on_host.py, in the GPU-less container) the two search entry points are this repo's KERNEL SOURCE compiled for the host
// (tests/host_emulation) and only level 1 runs.
#include "interpolation.hpp"
#include "interpolation_hd.hpp"
#include <cstring>
#include <random>

void ref_apply_inlet_outlet(LBM& lbm, const std::string& downstream_bc, const InletVelocityField& inlet, bool downstream_open_face, unsigned long min_work_per_thread, bool show_progress, int side_ref_z_cap_index);
void ref_apply_inlet_outlet_hd(LBM& lbm, const std::string& downstream_bc, const InletVelocityFieldHD& inlet, bool downstream_open_face, unsigned long min_work_per_thread, bool show_progress, int side_ref_z_cap_index);
void luw_inlet_eval_nearest(int device, const NearestNeighborInterpolator& nn, float z_threshold, const float3* pos, size_t count, float3* u);
void luw_inlet_eval_knn(int device, const KNNInterpolatorHD& knn, float z_base, const float3* pos, size_t count, float3* u);

namespace {

struct Cloud { const char* name; std::vector<float3> P, U; };

// samples on the five open faces of the box [-hx, hx] x [-hy, hy] x [-hz, hz] (the lattice's position() range), `per_face[f]` of them on face f
Cloud make_cloud(const char* name, const float hx, const float hy, const float hz, const int kind, const uint seed) {
	Cloud c; c.name = name;
	std::mt19937 rng(seed);
	std::uniform_real_distribution<float> unit(0.0f, 1.0f), vel(-0.1f, 0.1f);
	const auto add = [&](const float x, const float y, const float z) { c.P.push_back(float3(x, y, z)); c.U.push_back(float3(0.05f+vel(rng), vel(rng), 0.2f*vel(rng))); };
	const auto face_point = [&](const int f, const float a, const float b) { // a, b in [0, 1]
		if(f==0) add(-hx, -hy+2.0f*hy*a, -hz+2.0f*hz*b); else if(f==1) add(hx, -hy+2.0f*hy*a, -hz+2.0f*hz*b);
		else if(f==2) add(-hx+2.0f*hx*a, -hy, -hz+2.0f*hz*b); else if(f==3) add(-hx+2.0f*hx*a, hy, -hz+2.0f*hz*b);
		else add(-hx+2.0f*hx*a, -hy+2.0f*hy*b, hz);
	};
	if(kind==0) { // regular grids at a spacing of 2 cells through cell centres: equal distances everywhere, samples that coincide with cells
		for(int f=0; f<5; f++) {
			const float ha = f<2 ? hy : hx, hb = f<4 ? hz : hy;
			for(float a=-ha+0.5f; a<ha; a+=2.0f) for(float b=-hb+0.5f; b<hb; b+=2.0f) {
				if(f==0) add(-hx+0.5f, a, b); else if(f==1) add(hx-0.5f, a, b); else if(f==2) add(a, -hy+0.5f, b); else if(f==3) add(a, hy-0.5f, b); else add(a, b, hz-0.5f);
			}
		}
	} else if(kind==1) { // jittered clouds, a few hundred samples per face
		for(int f=0; f<5; f++) for(int i=0; i<300; i++) face_point(f, unit(rng), unit(rng));
	} else if(kind==2) { // sparse: 0 / 3 / 40 / 70 / 5 samples on the faces
		const int per_face[5] = { 0, 3, 40, 70, 5 };
		for(int f=0; f<5; f++) for(int i=0; i<per_face[f]; i++) face_point(f, unit(rng), unit(rng));
		add(-hx, 0.0f, hz); add(hx, 0.0f, hz); // keeps the bounding box on the faces
	} else if(kind==3) { // collinear samples on every face: the quadratic fit is singular
		for(int f=0; f<5; f++) for(int i=0; i<80; i++) face_point(f, 0.0125f*(float)i, 0.5f);
	} // kind 4: no samples
	return c;
}
std::vector<float3> face_positions(const uint Nx, const uint Ny, const uint Nz) { // the positions a lattice of that size evaluates: all five faces, z > 0
	std::vector<float3> p;
	const auto pos = [&](const uint x, const uint y, const uint z) { return float3((float)x-0.5f*(float)Nx+0.5f, (float)y-0.5f*(float)Ny+0.5f, (float)z-0.5f*(float)Nz+0.5f); };
	for(uint z=1u; z<Nz; z++) for(uint y=0u; y<Ny; y++) for(uint x=0u; x<Nx; x++) if(x==0u||x==Nx-1u||y==0u||y==Ny-1u||z==Nz-1u) p.push_back(pos(x, y, z));
	return p;
}
ulong differing(const std::vector<float3>& a, const std::vector<float3>& b) {
	ulong d = 0ull;
	for(size_t i=0u; i<a.size(); i++) { const float x[3] = { a[i].x, a[i].y, a[i].z }, y[3] = { b[i].x, b[i].y, b[i].z }; if(memcmp(x, y, sizeof(x))!=0) d++; }
	return d;
}

} // namespace

int luw_inlet_parity_main(const int device, const bool with_lattices) {
	int bad = 0, runs = 0;
	if(const char* golden = getenv("LUW_INLET_GOLDEN")) { // tests/golden/make_golden_inlet.py: the REFERENCE's velocities for small clouds, as a fixture for oracle/inlet_oracle.py
		FILE* f = fopen(golden, "wb");
		if(!f) return 1;
		const uint Gx = 23u, Gy = 19u, Gz = 13u;
		const std::vector<float3> gp = face_positions(Gx, Gy, Gz);
		const uint kinds[4] = { 0u, 1u, 2u, 3u };
		const uint head[2] = { 4u, (uint)gp.size() };
		fwrite(head, sizeof(uint), 2u, f);
		fwrite(gp.data(), sizeof(float3), gp.size(), f);
		for(int k=0; k<4; k++) {
			const Cloud c = make_cloud("golden", 0.5f*(float)Gx, 0.5f*(float)Gy, 0.5f*(float)Gz, (int)kinds[k], 91u+(uint)k);
			const float z_threshold = k%2 ? -3.25f : -1.0E9f;
			std::vector<float3> hd(gp.size()), lo(gp.size());
			KNNInterpolatorHD knn(c.P, c.U); InletVelocityFieldHD fhd(knn, z_threshold);
			NearestNeighborInterpolator nn(c.P, c.U); InletVelocityField flo(nn, z_threshold, 0.0f);
			for(size_t i=0u; i<gp.size(); i++) { hd[i] = fhd(gp[i]); lo[i] = flo(gp[i]); }
			const uint n = (uint)c.P.size();
			fwrite(&n, sizeof(uint), 1u, f); fwrite(&z_threshold, sizeof(float), 1u, f);
			fwrite(c.P.data(), sizeof(float3), c.P.size(), f); fwrite(c.U.data(), sizeof(float3), c.U.size(), f);
			fwrite(hd.data(), sizeof(float3), hd.size(), f); fwrite(lo.data(), sizeof(float3), lo.size(), f);
		}
		fclose(f);
		printf("golden inlet fixture: %u positions x 4 clouds -> %s\n", (uint)gp.size(), golden);
		return 0;
	}
	if(getenv("LUW_INLET_TIMING")) { // how long the reference's own per-cell evaluation takes on one host thread: 100 000 samples (20 000 per face), 512 positions
		Cloud c; c.name = "timing";
		std::mt19937 rng(1u); std::uniform_real_distribution<float> unit(-512.0f, 512.0f);
		for(int f=0; f<5; f++) for(int i=0; i<20000; i++) { const float a = unit(rng), b = 0.25f*unit(rng); c.P.push_back(f==0 ? float3(-512.0f, a, b) : f==1 ? float3(512.0f, a, b) : f==2 ? float3(a, -512.0f, b) : f==3 ? float3(a, 512.0f, b) : float3(a, unit(rng), 128.0f)); c.U.push_back(float3(0.05f)); }
		std::vector<float3> p; for(int i=0; i<512; i++) p.push_back(float3(unit(rng), unit(rng), 127.5f));
		KNNInterpolatorHD knn(c.P, c.U); InletVelocityFieldHD hd(knn, -1.0E9f);
		NearestNeighborInterpolator nn(c.P, c.U); InletVelocityField lo(nn, -1.0E9f, 0.0f);
		float sink = 0.0f;
		Clock clock; for(const float3& q : p) sink += hd(q).x; const double t_hd = clock.stop();
		clock.start(); for(const float3& q : p) sink += lo(q).x; const double t_lo = clock.stop();
		printf("reference evaluation on one host thread, 100000 samples: KNN-HD %.1f us per cell, nearest %.1f us per cell (%g)\n", 1.0E6*t_hd/512.0, 1.0E6*t_lo/512.0, sink);
		// the host part of this repo's K = 64 path (the fit over the kept samples): 64 samples on the top face, so that the search is trivial; LBM_NUM_THREADS=1 for a per-thread figure
		Cloud small; small.name = "fit";
		for(int i=0; i<64; i++) { small.P.push_back(float3(unit(rng), unit(rng), 128.0f)); small.U.push_back(float3(0.05f, 0.01f, 0.0f)); }
		small.P.push_back(float3(-512.0f, -512.0f, -64.0f)); small.U.push_back(float3(0.0f)); small.P.push_back(float3(512.0f, 512.0f, -64.0f)); small.U.push_back(float3(0.0f)); // the cloud's bounding box
		std::vector<float3> many, out(200000); for(int i=0; i<200000; i++) many.push_back(float3(0.9f*unit(rng), 0.9f*unit(rng), 127.5f));
		KNNInterpolatorHD knn64(small.P, small.U);
		clock.start(); luw_inlet_eval_knn(device, knn64, -1.0E9f, many.data(), many.size(), out.data()); const double t_fit = clock.stop();
		printf("this repo, 200000 cells x 64 samples (search trivial): %.2f us per cell wall clock with LBM_NUM_THREADS=%s\n", 1.0E6*t_fit/200000.0, getenv("LBM_NUM_THREADS") ? getenv("LBM_NUM_THREADS") : "(all)");
		return 0;
	}
	const uint Nx = 41u, Ny = 34u, Nz = 27u;
	const std::vector<float3> pos = face_positions(Nx, Ny, Nz);
	const char* kinds[5] = { "regular grid", "jittered", "sparse faces", "collinear", "no samples" };
	for(int kind=0; kind<5; kind++) for(int thr=0; thr<2; thr++) {
		const Cloud c = make_cloud(kinds[kind], 0.5f*(float)Nx, 0.5f*(float)Ny, 0.5f*(float)Nz, kind, 7u+(uint)kind);
		const float z_threshold = thr ? -4.25f : -1.0E9f;
		std::vector<float3> ref(pos.size()), ours(pos.size());
		{ // nearest sample
			NearestNeighborInterpolator nn(c.P, c.U);
			InletVelocityField field(nn, z_threshold, 0.0f);
			for(size_t i=0u; i<pos.size(); i++) ref[i] = field(pos[i]);
			luw_inlet_eval_nearest(device, nn, z_threshold+0.0f, pos.data(), pos.size(), ours.data());
			const ulong d = differing(ref, ours);
			printf("inlet parity, positions: nearest %-13s threshold %d: %s (%llu of %llu velocities differ)\n", c.name, thr, d==0ull ? "IDENTICAL" : "DIFFERENT", (unsigned long long)d, (unsigned long long)pos.size());
			bad += d==0ull ? 0 : 1; runs++;
		}
		{ // K = 64 quadratic fit
			KNNInterpolatorHD knn(c.P, c.U);
			InletVelocityFieldHD field(knn, z_threshold);
			for(size_t i=0u; i<pos.size(); i++) ref[i] = field(pos[i]);
			luw_inlet_eval_knn(device, knn, z_threshold, pos.data(), pos.size(), ours.data());
			const ulong d = differing(ref, ours);
			ulong nonzero = 0ull; for(const float3& v : ref) if(v.x!=0.0f||v.y!=0.0f||v.z!=0.0f) nonzero++;
			printf("inlet parity, positions: KNN-HD  %-13s threshold %d: %s (%llu of %llu velocities differ, %llu non-zero)\n", c.name, thr, d==0ull ? "IDENTICAL" : "DIFFERENT", (unsigned long long)d, (unsigned long long)pos.size(), (unsigned long long)nonzero);
			bad += d==0ull ? 0 : 1; runs++;
		}
	}
#ifndef LUW_INLET_ON_HOST
	if(with_lattices&&!getenv("LUW_INLET_PARITY_POSITIONS_ONLY")) { // (the sanitizer runs use the position level only)
		const uint shapes[3][3] = { { 37u, 29u, 23u }, { 64u, 3u, 17u }, { 5u, 41u, 9u } };
		const char* downs[5] = { "+x", "-x", "+y", "-y", "none" };
		for(int s=0; s<3; s++) for(int d=0; d<5; d++) for(int open=0; open<2; open++) for(int hd=0; hd<2; hd++) {
			const int cap = (d+open)%2 ? (int)shapes[s][2]/2 : -1;
			const Cloud c = make_cloud("jittered", 0.5f*(float)shapes[s][0], 0.5f*(float)shapes[s][1], 0.5f*(float)shapes[s][2], (s+d)%2, 31u+(uint)s);
			LBM a(uint3(shapes[s][0], shapes[s][1], shapes[s][2]), 1u, 1u, 1u, 0.01f), b(uint3(shapes[s][0], shapes[s][1], shapes[s][2]), 1u, 1u, 1u, 0.01f);
			std::mt19937 rng(5u+(uint)s);
			for(ulong n=0ull; n<a.get_N(); n++) { // what the voxeliser and earlier set-up steps may have left behind
				const uchar f = (uchar)(rng()%9u==0u ? TYPE_S : 0u); const float v = 1.0E-3f*(float)(rng()%100u);
				a.flags[n] = f; b.flags[n] = f; a.u.x[n] = v; b.u.x[n] = v; a.u.y[n] = -v; b.u.y[n] = -v; a.u.z[n] = 0.5f*v; b.u.z[n] = 0.5f*v;
			}
			const float z_threshold = -0.25f*(float)shapes[s][2];
			if(hd) {
				KNNInterpolatorHD knn(c.P, c.U); InletVelocityFieldHD field(knn, z_threshold);
				ref_apply_inlet_outlet_hd(a, downs[d], field, open!=0, 500000ull, false, cap);
				apply_inlet_outlet_hd(b, downs[d], field, open!=0, 500000ull, false, cap);
			} else {
				NearestNeighborInterpolator nn(c.P, c.U); InletVelocityField field(nn, z_threshold, 0.5f);
				ref_apply_inlet_outlet(a, downs[d], field, open!=0, 500000ull, false, cap);
				apply_inlet_outlet(b, downs[d], field, open!=0, 500000ull, false, cap);
			}
			ulong diff = 0ull, type_e = 0ull;
			for(ulong n=0ull; n<a.get_N(); n++) {
				const float ua[3] = { a.u.x[n], a.u.y[n], a.u.z[n] }, ub[3] = { b.u.x[n], b.u.y[n], b.u.z[n] };
				if(a.flags[n]!=b.flags[n]||memcmp(ua, ub, sizeof(ua))!=0) diff++;
				if(a.flags[n]==TYPE_E) type_e++;
			}
			printf("inlet parity, lattice %ux%ux%u %s downstream %-4s open %d cap %d: %s (cells differing %llu, TYPE_E %llu)\n", shapes[s][0], shapes[s][1], shapes[s][2], hd ? "KNN-HD " : "nearest", downs[d], open, cap,
				diff==0ull ? "IDENTICAL" : "DIFFERENT", (unsigned long long)diff, (unsigned long long)type_e);
			bad += diff==0ull ? 0 : 1; runs++;
		}
	}
#else
	(void)with_lattices;
#endif
	uint64_t launches = 0ull;
	luw_inlet_launch_count(&launches);
	printf("inlet parity: %d of %d runs identical; search kernels launched: %llu\n", runs-bad, runs, (unsigned long long)launches);
	return bad;
}
