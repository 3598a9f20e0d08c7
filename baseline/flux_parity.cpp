// TEST HARNESS (baseline/build_reference_driver.py links it into baseline/_ref/luw_flux_parity; tests/test_reference_driver.py runs it on the GPU box):
// the reference's own apply_flux_correction (FX/fluxcorrection.cpp, compiled unmodified under the name ref_apply_flux_correction) and this repo's O(surface)
// replacement (latticeurbanwind_b200/host/fluxcorrection_surface.cpp) on two LBM objects with identical host images -> flags, u and the reported sums must be
// bit-identical, for every downstream face, with and without the inflow re-evaluation.
#include "fluxcorrection.hpp"
#include <cstring>
#include <random>

void ref_apply_flux_correction(LBM& lbm, const std::string& downstream_bc, const std::function<float3(const float3&)>& inlet_eval, bool show_report,
	double* avg_delta_mps, double* net_before, double* net_after);

static void fill(LBM& lbm, const uint seed) {
	std::mt19937 rng(seed);
	std::uniform_real_distribution<float> vel(-0.12f, 0.12f);
	for(ulong n=0ull; n<lbm.get_N(); n++) {
		const uint r = rng()%100u;
		lbm.flags[n] = (uchar)((r<12u ? TYPE_S : r<18u ? TYPE_E : 0u)|(rng()%7u==0u ? TYPE_T : 0u));
		lbm.u.x[n] = vel(rng); lbm.u.y[n] = vel(rng); lbm.u.z[n] = vel(rng);
	}
}

int reference_main_unused();
int luw_flux_parity_main() {
	int bad = 0, runs = 0;
	const uint shapes[3][3] = { { 37u, 29u, 23u }, { 64u, 3u, 17u }, { 5u, 41u, 9u } };
	const char* downs[5] = { "+x", "-x", "+y", "-y", "none" };
	for(int s=0; s<3; s++) for(int d=0; d<5; d++) for(int refill=0; refill<2; refill++) {
		LBM a(uint3(shapes[s][0], shapes[s][1], shapes[s][2]), 1u, 1u, 1u, 0.01f), b(uint3(shapes[s][0], shapes[s][1], shapes[s][2]), 1u, 1u, 1u, 0.01f);
		fill(a, 100u+(uint)s); fill(b, 100u+(uint)s);
		const std::function<float3(const float3&)> eval = refill ? std::function<float3(const float3&)>([](const float3& p) { return float3(0.05f+1.0E-3f*p.z, -2.0E-3f*p.x, 1.0E-4f*p.y); }) : std::function<float3(const float3&)>();
		double ra[3] = { 0.0, 0.0, 0.0 }, rb[3] = { 0.0, 0.0, 0.0 };
		ref_apply_flux_correction(a, downs[d], eval, false, &ra[0], &ra[1], &ra[2]);
		apply_flux_correction(b, downs[d], eval, false, &rb[0], &rb[1], &rb[2]);
		ulong diff = 0ull;
		for(ulong n=0ull; n<a.get_N(); n++) {
			const float ua[3] = { a.u.x[n], a.u.y[n], a.u.z[n] }, ub[3] = { b.u.x[n], b.u.y[n], b.u.z[n] };
			if(a.flags[n]!=b.flags[n]||memcmp(ua, ub, sizeof(ua))!=0) diff++;
		}
		const bool same = diff==0ull&&memcmp(ra, rb, sizeof(ra))==0;
		printf("flux parity %ux%ux%u downstream %-4s refill %d: %s (cells differing %llu; net_before %.17g / %.17g, net_after %.3e / %.3e)\n", shapes[s][0], shapes[s][1], shapes[s][2], downs[d], refill,
			same ? "IDENTICAL" : "DIFFERENT", (unsigned long long)diff, ra[1], rb[1], ra[2], rb[2]);
		bad += same ? 0 : 1; runs++;
	}
	printf("flux parity: %d of %d runs identical\n", runs-bad, runs);
	return bad;
}
