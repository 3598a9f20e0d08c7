// TEST INFRASTRUCTURE (see cuda_on_host.hpp): the cell arithmetic of csrc/lbm_vec.cuh compiled for the host. The FAST two-pass formulation (moments_of /
// fast_prepare / fast_relax_*: an algebraic regrouping of the collision) is evaluated next to the as-written STRICT formulation (collide_strict2, which the GPU
// tests pin bit for bit to the oracle) on the same DDFs, so that a wrong coefficient or sign in the regrouping shows up in the container that has no GPU.
#define LUW_HOST_EMULATION 1
#include "cuda_on_host.hpp"
#include "../../latticeurbanwind_b200/csrc/lbm_vec.cuh"
using namespace luw;

template<uint32_t FEAT> static void one_pair(const DomainConst& c, const StepArgs& a, const float* f0, const float* f1, const float scale, float* strict, float* fast, float* ru_strict, float* ru_fast) {
	PairIn in; in.zones = false; in.e0 = in.e1 = false; in.any_e = false;
	f2 f[Q], g[Q];
	for(int i=0; i<Q; i++) { f[i] = mk2(f0[i], f1[i]); g[i] = mk2(scale*f0[i], scale*f1[i]); }
	PairOut os, of;
	collide_strict2<FEAT>(c, a, in, f, os);
	for(int i=0; i<Q; i++) { strict[i] = f[i].v.x; strict[Q+i] = f[i].v.y; }
	Moments M;
	const auto ld = [&](const int k, f2& gi, f2& gj) { gi = g[2*k+1]; gj = g[2*k+2]; };
	moments_of<(FEAT&F_SUBGRID)!=0u>(g[0], ld, M);
	FastK K;
	fast_prepare<FEAT>(c, a, in, M, scale, 1.0f/scale, K, of);
	g[0] = fma2(K.omw, g[0], K.g0add);
	for(int ax=0; ax<3; ax++) fast_relax_axis(K, ax, g[2*ax+1], g[2*ax+2]);
	for(int pl=0; pl<3; pl++) fast_relax_diag(K, pl, g[2*(3+pl)+1], g[2*(3+pl)+2], g[2*(6+pl)+1], g[2*(6+pl)+2]);
	for(int i=0; i<Q; i++) { fast[i] = g[i].v.x/scale; fast[Q+i] = g[i].v.y/scale; }
	const f2 rs[4] = { os.rho, os.ux, os.uy, os.uz }, rf[4] = { of.rho, of.ux, of.uy, of.uz };
	for(int j=0; j<4; j++) { ru_strict[j] = rs[j].v.x; ru_strict[4+j] = rs[j].v.y; ru_fast[j] = rf[j].v.x; ru_fast[4+j] = rf[j].v.y; }
}

// TYPE_E lanes of the FAST two-pass path: rho / u from the boundary fields, relaxation rate 1, no forcing term -> the streamed-in DDFs (garbage here) are wiped and
// U +- V must be the equilibrium of (rho, u after the force half-step and the clamp). bnd: rho, ux, uy, uz of the two cells; out: the 2 x 19 post-"collision" DDFs / scale.
template<uint32_t FEAT> static void one_pair_e(const DomainConst& c, const StepArgs& a, const float* f0, const float* f1, const float* bnd, const float scale, float* fast) {
	PairIn in; in.zones = false; in.e0 = in.e1 = true; in.any_e = true; in.n = 0ull;
	float rho2[2] = { bnd[0], bnd[4] }, u6[6] = { bnd[1], bnd[5], bnd[2], bnd[6], bnd[3], bnd[7] }; // a two-cell "lattice": the boundary fields fast_prepare reads (c.N = 2)
	DomainConst cc = c; cc.rho = rho2; cc.u = u6; cc.N = 2ull;
	f2 g[Q];
	for(int i=0; i<Q; i++) g[i] = mk2(scale*f0[i], scale*f1[i]);
	Moments M;
	const auto ld = [&](const int k, f2& gi, f2& gj) { gi = g[2*k+1]; gj = g[2*k+2]; };
	moments_of<(FEAT&F_SUBGRID)!=0u>(g[0], ld, M);
	FastK K; PairOut of;
	fast_prepare<FEAT>(cc, a, in, M, scale, 1.0f/scale, K, of);
	g[0] = fma2(K.omw, g[0], K.g0add);
	for(int ax=0; ax<3; ax++) fast_relax_axis(K, ax, g[2*ax+1], g[2*ax+2]);
	for(int pl=0; pl<3; pl++) fast_relax_diag(K, pl, g[2*(3+pl)+1], g[2*(3+pl)+2], g[2*(6+pl)+1], g[2*(6+pl)+2]);
	for(int i=0; i<Q; i++) { fast[i] = g[i].v.x/scale; fast[Q+i] = g[i].v.y/scale; }
}
extern "C" int emu_fast_equilibrium(uint32_t feat, uint64_t npairs, const float* f, const float* bnd, float w, const float* force_omega6, float scale, float* fast) {
	DomainConst c; memset(&c, 0, sizeof(c));
	c.w = w; c.tau0 = 1.0f/w; c.tau0sq = c.tau0*c.tau0; c.features = feat;
	const StepArgs a = { 0ull, force_omega6[0], force_omega6[1], force_omega6[2], force_omega6[3], force_omega6[4], force_omega6[5] };
	for(uint64_t p=0; p<npairs; p++) {
		const float* f0 = f+p*2*Q; const float* f1 = f0+Q;
		switch(feat&15u) {
			case 4u: one_pair_e<4u>(c, a, f0, f1, bnd+p*8, scale, fast+p*2*Q); break;
			case 14u: one_pair_e<14u>(c, a, f0, f1, bnd+p*8, scale, fast+p*2*Q); break;
			case 15u: one_pair_e<15u>(c, a, f0, f1, bnd+p*8, scale, fast+p*2*Q); break;
			default: return 1;
		}
	}
	return 0;
}

extern "C" int emu_fast_vs_strict(uint32_t feat, uint64_t npairs, const float* f, float w, const float* force_omega6, float scale, float* strict, float* fast, float* ru_strict, float* ru_fast) {
	DomainConst c; memset(&c, 0, sizeof(c));
	c.w = w; c.tau0 = 1.0f/w; c.tau0sq = c.tau0*c.tau0; c.features = feat;
	const StepArgs a = { 0ull, force_omega6[0], force_omega6[1], force_omega6[2], force_omega6[3], force_omega6[4], force_omega6[5] };
	for(uint64_t p=0; p<npairs; p++) {
		const float* f0 = f+p*2*Q; const float* f1 = f0+Q;
		float* s = strict+p*2*Q; float* g = fast+p*2*Q; float* rs = ru_strict+p*8; float* rf = ru_fast+p*8;
		switch(feat&15u) {
			case 0u: one_pair<0u>(c, a, f0, f1, scale, s, g, rs, rf); break;
			case 4u: one_pair<4u>(c, a, f0, f1, scale, s, g, rs, rf); break;
			case 5u: one_pair<5u>(c, a, f0, f1, scale, s, g, rs, rf); break;
			case 6u: one_pair<6u>(c, a, f0, f1, scale, s, g, rs, rf); break;
			case 12u: one_pair<12u>(c, a, f0, f1, scale, s, g, rs, rf); break;
			case 14u: one_pair<14u>(c, a, f0, f1, scale, s, g, rs, rf); break;
			case 15u: one_pair<15u>(c, a, f0, f1, scale, s, g, rs, rf); break;
			default: return 1;
		}
	}
	return 0;
}
