// End-to-end timing THROUGH THE REFERENCE'S C++ API (bench.py's e2e.cpp_host leg): an LBM object built and driven the way FX/setup.cpp drives it -- boundary
// conditions written through lbm.flags / lbm.rho / lbm.u, run(0) to initialise, then one run(1) per time step (FX/setup.cpp:4898) -- with HOST buffers: every timed
// step uploads the velocity of the top boundary plane from the host mirror and reads rho / u of a probe plane back into it (LBM_Domain::u.enqueue_write_to_device /
// enqueue_read_from_device(offset, length), FX/opencl.hpp:481-512), synchronising once per step like FX/lbm.cpp:1288.
//   luw_host_bench <channel|urban|rest> Nx Ny Nz precision features steps warmup
//   rest: nothing is written on the host (fluid at rest, periodic): the lazy host mirrors stay untouched -- the >= 1 G-cell case of tests/test_cpp_host.py
// Prints one JSON line.
#include "lbm.hpp"
#include <chrono>
#include <cmath>
#include <sys/resource.h>

static double now() { return std::chrono::duration_cast<std::chrono::duration<double>>(std::chrono::high_resolution_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
	if(argc!=9) { fprintf(stderr, "usage: luw_host_bench <channel|urban|rest> Nx Ny Nz precision features steps warmup\n"); return 2; }
	const std::string kind = argv[1];
	const uint Nx = (uint)atoi(argv[2]), Ny = (uint)atoi(argv[3]), Nz = (uint)atoi(argv[4]);
	lbm_settings.precision = (uint)atoi(argv[5]);
	const uint want = (uint)atoi(argv[6]);
	lbm_settings.features = want&~(uint)(LUW_BUFFER_NUDGING|LUW_TOP_SPONGE);
	lbm_settings.arith = LUW_ARITH_FAST;
	lbm_settings.downstream_face = 2;
	if(want&LUW_BUFFER_NUDGING) lbm_settings.set_buffer_nudging(16u, 0.01f, true);
	if(want&LUW_TOP_SPONGE) lbm_settings.set_top_sponge(20u, 0.02f);
	const ulong steps = (ulong)atoll(argv[7]), warmup = (ulong)atoll(argv[8]);
	const float nu = kind=="urban" ? 1.0E-6f : 1.0f/6.0f;
	const double t_build0 = now();
	LBM lbm(uint3(Nx, Ny, Nz), 1u, 1u, 1u, nu);
	if(want&LUW_VOLUME_FORCE) lbm.set_coriolis(0.0f, 5.6E-6f, 4.7E-6f);
	const ulong N = lbm.get_N(), plane = (ulong)Nx*(ulong)Ny;
	if(kind!="rest") { // the case, through the stitched global accessors like FX/setup.cpp:4931-5353
		for(ulong n=0ull; n<N; n++) {
			uint x, y, z; lbm.coordinates(n, x, y, z);
			uchar fl = 0u; float ux = kind=="urban" ? 0.1f*logf(1.0f+(float)z)/logf(1.0f+(float)Nz) : 0.05f;
			if(kind=="channel") {
				if(y==0u||y==Ny-1u||z==0u||z==Nz-1u) fl = TYPE_S; else if(x==0u||x==Nx-1u) fl = TYPE_E;
			} else { // staggered cube array (pitch 32, edge 16, heights 16 + 8 k) over the central 3/4 of the ground plane, open sides and top
				const uint x0 = Nx/8u, x1 = Nx-Nx/8u, y0 = Ny/8u, y1 = Ny-Ny/8u;
				if(z==0u) fl = TYPE_S;
				else if(x==0u||x==Nx-1u||y==0u||y==Ny-1u||z==Nz-1u) fl = TYPE_E;
				else if(x>=x0&&x<x1&&y>=y0&&y<y1) {
					const uint iy = (y-y0)/32u, xs = x-x0, sh = (iy%2u)*16u;
					if(xs>=sh&&(xs-sh)%32u<16u&&(y-y0)%32u<16u&&z<=16u+8u*((((xs-sh)/32u)*7u+iy*13u)%5u)) fl = TYPE_S;
				}
			}
			if(fl==TYPE_S) ux = 0.0f;
			lbm.flags[n] = fl; lbm.u.x[n] = ux;
		}
	}
	lbm.run(0ull);
	const double t_build = now()-t_build0;
	LBM_Domain* dom = lbm.lbm_domain[0];
	const auto step = [&](const bool io) {
		if(io) for(uint c=0u; c<3u; c++) dom->u.enqueue_write_to_device((ulong)c*N+(ulong)(Nz-1u)*plane, plane); // this step's top-boundary velocity (TYPE_E plane), 12 B per face cell
		lbm.run(1ull);
		if(io) {
			for(uint c=0u; c<3u; c++) dom->u.enqueue_read_from_device((ulong)c*N+(ulong)(Nz/2u)*plane, plane); // probe plane z = Nz/2
			dom->rho.enqueue_read_from_device((ulong)(Nz/2u)*plane, plane);
			dom->finish_queue();
		}
	};
	const bool io = kind!="rest";
	for(ulong k=0ull; k<warmup; k++) step(io);
	dom->finish_queue();
	const double t0 = now();
	for(ulong k=0ull; k<steps; k++) step(io);
	dom->finish_queue();
	const double dt = now()-t0;
	double probe = 0.0;
	if(io) { for(ulong i=0ull; i<plane; i++) probe += (double)dom->u[(ulong)(Nz/2u)*plane+i]; probe /= (double)plane; }
	else { dom->u.read_from_device(N/2ull, 4096ull); for(ulong i=0ull; i<4096ull; i++) probe += fabs((double)dom->u[N/2ull+i]); } // a fluid at rest stays at rest; touches 16 KB of a pageable mirror
	struct rusage ru; getrusage(RUSAGE_SELF, &ru);
	printf("{\"case\": \"%s\", \"lattice\": [%u, %u, %u], \"cells\": %llu, \"steps\": %llu, \"ms_per_step\": %.6f, \"mlups\": %.1f, \"h2d_bytes_per_step\": %llu, \"d2h_bytes_per_step\": %llu, "
		"\"probe_mean_ux\": %.9g, \"build_and_initialize_s\": %.3f, \"tiled\": %d, \"host_mirror_rho\": %d, \"host_mirror_u\": %d, \"host_mirror_flags\": %d, \"peak_rss_mb\": %ld, \"device_mb\": %llu}\n",
		kind.c_str(), Nx, Ny, Nz, (unsigned long long)N, (unsigned long long)steps, dt/(double)steps*1.0E3, (double)N*(double)steps/dt/1.0E6,
		(unsigned long long)(io ? 12ull*plane : 0ull), (unsigned long long)(io ? 16ull*plane : 0ull), probe, t_build, (int)dom->uses_tiles(),
		(int)dom->rho.materialized(), (int)dom->u.materialized(), (int)dom->flags.materialized(), ru.ru_maxrss/1024l, (unsigned long long)(dom->device_memory_used()/1048576ull));
	return 0;
}
