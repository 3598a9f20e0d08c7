// Cell arithmetic of the LBM step on sm_100a's packed FP32 datapath.
//
// Blackwell issues FADD2 / FMUL2 / FFMA2: two IEEE-rounded FP32 operations per lane per instruction (64-bit register pairs). The
// step kernel updates TWO x-adjacent cells per thread, so every add / multiply / fma of the collision is one packed instruction for
// both cells. `f2` wraps that; each lane is rounded exactly like the scalar operation, so the STRICT formulation below (operations in
// the reference's order, fused only where the reference writes fma(), FX/kernel.cpp:1016-1113,1686-1748) is bit-identical to the oracle.
// The FAST formulation is an algebraically regrouped collision (pair sums / differences of opposite directions shared between
// moments, equilibrium, Smagorinsky tensor and Guo forcing; approximate reciprocal and square root) for the tolerance-checked mode.
//
// All operations are explicit intrinsics, so the result does not depend on the translation unit's -fmad setting.
#pragma once
#include "lbm_common.cuh"

namespace luw {
namespace {

struct f2 { float2 v; };
__device__ __forceinline__ f2 mk2(const float a, const float b) { f2 r; r.v = make_float2(a, b); return r; }
__device__ __forceinline__ f2 bc(const float a) { return mk2(a, a); }
__device__ __forceinline__ f2 operator+(const f2 a, const f2 b) { f2 r; r.v = __fadd2_rn(a.v, b.v); return r; }
__device__ __forceinline__ f2 operator-(const f2 a, const f2 b) { f2 r; r.v = __fadd2_rn(a.v, make_float2(-b.v.x, -b.v.y)); return r; }
__device__ __forceinline__ f2 operator*(const f2 a, const f2 b) { f2 r; r.v = __fmul2_rn(a.v, b.v); return r; }
__device__ __forceinline__ f2 operator*(const float a, const f2 b) { f2 r; r.v = __fmul2_rn(make_float2(a, a), b.v); return r; }
__device__ __forceinline__ f2 operator-(const f2 a) { return mk2(-a.v.x, -a.v.y); }
__device__ __forceinline__ f2 fma2(const f2 a, const f2 b, const f2 c) { f2 r; r.v = __ffma2_rn(a.v, b.v, c.v); return r; }
__device__ __forceinline__ f2 fma2(const float a, const f2 b, const f2 c) { f2 r; r.v = __ffma2_rn(make_float2(a, a), b.v, c.v); return r; }
__device__ __forceinline__ f2 fma2(const f2 a, const float b, const f2 c) { f2 r; r.v = __ffma2_rn(a.v, make_float2(b, b), c.v); return r; }
__device__ __forceinline__ f2 fma2(const f2 a, const f2 b, const float c) { f2 r; r.v = __ffma2_rn(a.v, b.v, make_float2(c, c)); return r; }
__device__ __forceinline__ f2 div_rn2(const f2 a, const f2 b) { return mk2(__fdiv_rn(a.v.x, b.v.x), __fdiv_rn(a.v.y, b.v.y)); }
__device__ __forceinline__ f2 div_rn2(const float a, const f2 b) { return mk2(__fdiv_rn(a, b.v.x), __fdiv_rn(a, b.v.y)); }
__device__ __forceinline__ f2 sqrt_rn2(const f2 a) { return mk2(__fsqrt_rn(a.v.x), __fsqrt_rn(a.v.y)); }
__device__ __forceinline__ f2 clampc2(const f2 a) { return mk2(fminf(fmaxf(a.v.x, -LAT_C), LAT_C), fminf(fmaxf(a.v.y, -LAT_C), LAT_C)); }
#ifndef LUW_HOST_EMULATION
__device__ __forceinline__ float rcp_approx(const float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sqrt_approx(const float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#else // tests/host_emulation: the algebra of the FAST formulation is checked on the host against the as-written one
__device__ __forceinline__ float rcp_approx(const float x) { return 1.0f/x; }
__device__ __forceinline__ float sqrt_approx(const float x) { return sqrtf(x); }
#endif
__device__ __forceinline__ f2 rcp2(const f2 a) { return mk2(rcp_approx(a.v.x), rcp_approx(a.v.y)); }
__device__ __forceinline__ f2 sqrt2(const f2 a) { return mk2(sqrt_approx(a.v.x), sqrt_approx(a.v.y)); }
__device__ __forceinline__ f2 sel2(const bool c0, const bool c1, const f2 a, const f2 b) { return mk2(c0 ? a.v.x : b.v.x, c1 ? a.v.y : b.v.y); }

// relaxation-zone data of one cell (FX/kernel.cpp:1523-1614): gathered BEFORE the DDFs are read so that the dependent global loads
// (reference velocity at the boundary face / top plane) are in flight while the moments are computed
struct ZoneRef {
	bool nudge, sponge; // cell lies in the buffer-nudging shell / in the top sponge (and is not TYPE_E)
	float kn, unx, uny, unz; // nudging: w_buf*inv_tau and the target velocity u[n_ref]
	float ks, usx, usy, usz; // sponge: sigma and the reference velocity at the top plane
};
__device__ __forceinline__ ZoneRef zone_prefetch(const DomainConst& c, const uint32_t x, const uint32_t y, const uint32_t z, const bool active) {
	ZoneRef r;
	r.nudge = false; r.sponge = false; r.kn = 0.0f; r.unx = r.uny = r.unz = 0.0f; r.ks = 0.0f; r.usx = r.usy = r.usz = 0.0f;
	if(!active) return r;
	const uint64_t row = c.Px, plane = (uint64_t)c.Px*c.Ny;
	if(c.features&F_NUDGING) {
		const int xg = (int)x+c.Ox, yg = (int)y+c.Oy, zg = (int)z+c.Oz, Nb = (int)c.buffer_N;
		const int dw = xg, de = (int)(c.Nxg-1u)-xg, ds = yg, dn = (int)(c.Nyg-1u)-yg, dt = (int)(c.Nzg-1u)-zg;
		const bool in_w = c.downstream_face!=1&&c.has_w&&dw>=0&&dw<=Nb;
		const bool in_e = c.downstream_face!=2&&c.has_e&&de>=0&&de<=Nb;
		const bool in_s = c.downstream_face!=3&&c.has_s&&ds>=0&&ds<=Nb;
		const bool in_n = c.downstream_face!=4&&c.has_n&&dn>=0&&dn<=Nb;
		const bool in_t = c.has_t&&dt>=0&&dt<=Nb;
		if(in_w||in_e||in_s||in_n||in_t) {
			uint32_t dmin = c.buffer_N+1u;
			uint64_t nref = 0ull;
			if(in_w&&(uint32_t)dw<dmin) { dmin = (uint32_t)dw; nref = (uint64_t)c.wx+y*row+z*plane; }
			if(in_e&&(uint32_t)de<dmin) { dmin = (uint32_t)de; nref = (uint64_t)c.ex+y*row+z*plane; }
			if(in_s&&(uint32_t)ds<dmin) { dmin = (uint32_t)ds; nref = x+(uint64_t)c.sy*row+z*plane; }
			if(in_n&&(uint32_t)dn<dmin) { dmin = (uint32_t)dn; nref = x+(uint64_t)c.ny*row+z*plane; }
			if(in_t&&(uint32_t)dt<dmin) { dmin = (uint32_t)dt; nref = x+y*row+(uint64_t)c.tz*plane; }
			r.nudge = true;
			r.kn = __fmul_rn(__ldg(c.wbuf+dmin), c.buffer_inv_tau);
			r.unx = __ldg(c.u+nref); r.uny = __ldg(c.u+c.N+nref); r.unz = __ldg(c.u+2ull*c.N+nref);
		}
	}
	if((c.features&F_SPONGE)&&c.has_t) {
		const int dt = (int)(c.Nzg-2u)-((int)z+c.Oz);
		if(dt>=0&&dt<(int)c.sponge_N) {
			const uint64_t nref = x+y*row+(uint64_t)c.tz*plane;
			r.sponge = true;
			r.ks = __ldg(c.sigma+dt);
			r.usx = __ldg(c.u+nref); r.usy = __ldg(c.u+c.N+nref); r.usz = __ldg(c.u+2ull*c.N+nref);
		}
	}
	return r;
}
// the two contributions, added in the reference's order and with its roundings
__device__ __forceinline__ void zone_apply(const ZoneRef& r, const int nudge_vertical, const float rho, const float ux, const float uy, const float uz, float& fxn, float& fyn, float& fzn) {
	if(r.nudge) {
		const float ax = __fmul_rn(r.kn, __fadd_rn(r.unx, -ux));
		const float ay = __fmul_rn(r.kn, __fadd_rn(r.uny, -uy));
		const float az = nudge_vertical==1 ? __fmul_rn(r.kn, __fadd_rn(r.unz, -uz)) : 0.0f;
		fxn = __fadd_rn(fxn, __fmul_rn(rho, ax)); fyn = __fadd_rn(fyn, __fmul_rn(rho, ay)); fzn = __fadd_rn(fzn, __fmul_rn(rho, az));
	}
	if(r.sponge) {
		const float rs = __fmul_rn(rho, r.ks);
		fxn = __fadd_rn(fxn, __fmul_rn(rs, __fadd_rn(r.usx, -ux)));
		fyn = __fadd_rn(fyn, __fmul_rn(rs, __fadd_rn(r.usy, -uy)));
		fzn = __fadd_rn(fzn, __fmul_rn(rs, __fadd_rn(r.usz, -uz)));
	}
}

struct PairIn { // what the step needs to know about the two cells besides their DDFs
	uint64_t n; // FAST two-pass, any_e: device index of the pair's first cell (the TYPE_E lanes' rho / u are loaded where they are needed: no registers are held across the moment pass)
	bool any_e; // FAST two-pass: e0 || e1 -- or, where the caller can afford a vote, "some lane of the warp holds a TYPE_E cell" (a warp-uniform branch then keeps the selects out of the common path)
	bool e0, e1; // FAST only: lane holds a TYPE_E cell -> rho/u are the boundary values below and f := feq (STRICT redoes such lanes in scalar code, lbm_tile.cuh)
	f2 rho_e, ux_e, uy_e, uz_e;
	bool zones; // some lane lies in a relaxation zone
	int nudge_vertical;
	ZoneRef zr0, zr1;
};
struct PairOut { f2 rho, ux, uy, uz; f2 upx, upy, upz; }; // the rho/u the reference writes with UPDATE_FIELDS (for non-TYPE_E cells); up*: the velocity before the force half-step (thermal step, DomainConst::upre)

__device__ __forceinline__ void zone_force2(const PairIn& in, const f2 rho, const f2 ux, const f2 uy, const f2 uz, f2& Fx, f2& Fy, f2& Fz) {
	zone_apply(in.zr0, in.nudge_vertical, rho.v.x, ux.v.x, uy.v.x, uz.v.x, Fx.v.x, Fy.v.x, Fz.v.x);
	zone_apply(in.zr1, in.nudge_vertical, rho.v.y, ux.v.y, uy.v.y, uz.v.y, Fx.v.y, Fy.v.y, Fz.v.y);
}

// =================================================================== STRICT: the reference's operations, in its order
// ptxas fuses mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with -fmad=false (measured, CUDA 12.9); products that must be rounded on
// their own are therefore formed with scalar FMULs (sm), which are never contracted; sums and explicit fmas stay packed.
__device__ __forceinline__ f2 sm(const f2 a, const f2 b) { return mk2(__fmul_rn(a.v.x, b.v.x), __fmul_rn(a.v.y, b.v.y)); }
__device__ __forceinline__ f2 sm(const float a, const f2 b) { return mk2(__fmul_rn(a, b.v.x), __fmul_rn(a, b.v.y)); }

__device__ __forceinline__ void feq_strict2(const f2 rho, f2 ux, f2 uy, f2 uz, f2* feq) { // FX/kernel.cpp:1016-1073 (D3Q19 branch)
	const f2 rhom1 = rho-bc(1.0f);
	const f2 c3 = sm(-3.0f, (sm(ux, ux)+sm(uy, uy))+sm(uz, uz));
	uz = sm(3.0f, uz); ux = sm(3.0f, ux); uy = sm(3.0f, uy);
	feq[0] = sm(W0, fma2(rho, sm(0.5f, c3), rhom1));
	const f2 u0 = ux+uy, u1 = ux+uz, u2 = uy+uz, u3 = ux-uy, u4 = ux-uz, u5 = uy-uz;
	const f2 rhos = sm(WS, rho), rhoe = sm(WE, rho), rhom1s = sm(WS, rhom1), rhom1e = sm(WE, rhom1);
#define LUW_FEQ2(i, a, r, r1) { const f2 q_ = fma2(a, a, c3); feq[i] = fma2(r, fma2(0.5f, q_, a), r1); feq[i+1] = fma2(r, fma2(0.5f, q_, -(a)), r1); }
	LUW_FEQ2( 1, ux, rhos, rhom1s) LUW_FEQ2( 3, uy, rhos, rhom1s) LUW_FEQ2( 5, uz, rhos, rhom1s)
	LUW_FEQ2( 7, u0, rhoe, rhom1e) LUW_FEQ2( 9, u1, rhoe, rhom1e) LUW_FEQ2(11, u2, rhoe, rhom1e)
	LUW_FEQ2(13, u3, rhoe, rhom1e) LUW_FEQ2(15, u4, rhoe, rhom1e) LUW_FEQ2(17, u5, rhoe, rhom1e)
#undef LUW_FEQ2
}
__device__ __forceinline__ void rho_u_strict2(const f2* f, f2& rho, f2& ux, f2& uy, f2& uz) { // FX/kernel.cpp:1075-1100
	f2 r = f[0];
#pragma unroll
	for(int i=1; i<Q; i++) r = r+f[i];
	r = r+bc(1.0f);
	const f2 mx = ((((((((f[ 1]-f[ 2])+f[ 7])-f[ 8])+f[ 9])-f[10])+f[13])-f[14])+f[15])-f[16];
	const f2 my = ((((((((f[ 3]-f[ 4])+f[ 7])-f[ 8])+f[11])-f[12])+f[14])-f[13])+f[17])-f[18];
	const f2 mz = ((((((((f[ 5]-f[ 6])+f[ 9])-f[10])+f[11])-f[12])+f[16])-f[15])+f[18])-f[17];
	rho = r; ux = div_rn2(mx, r); uy = div_rn2(my, r); uz = div_rn2(mz, r);
}
__device__ __forceinline__ void forcing_strict2(const f2 ux, const f2 uy, const f2 uz, const f2 fx, const f2 fy, const f2 fz, f2* Fin) { // FX/kernel.cpp:1103-1113
	const f2 uF = sm(-0.33333334f, fma2(ux, fx, fma2(uy, fy, sm(uz, fz))));
	const float t3 = 0.33333334f;
	Fin[0] = sm(9.0f*W0, uF);
	const float ks = 9.0f*WS, ke = 9.0f*WE;
#define LUW_FIN2(i, k, cf, cu) { const f2 cf_ = (cf), cu_ = (cu); Fin[i] = sm(k, fma2(cf_, cu_+bc(t3), uF)); Fin[i+1] = sm(k, fma2(-cf_, (-cu_)+bc(t3), uF)); }
	LUW_FIN2( 1, ks, fx, ux) LUW_FIN2( 3, ks, fy, uy) LUW_FIN2( 5, ks, fz, uz)
	LUW_FIN2( 7, ke, fx+fy, ux+uy) LUW_FIN2( 9, ke, fx+fz, ux+uz) LUW_FIN2(11, ke, fy+fz, uy+uz)
	LUW_FIN2(13, ke, fx-fy, ux-uy) LUW_FIN2(15, ke, fx-fz, ux-uz) LUW_FIN2(17, ke, fy-fz, uy-uz)
#undef LUW_FIN2
}
__device__ __forceinline__ f2 smagorinsky_strict2(const float w0, const f2* f, const f2* feq, const f2 rho) { // FX/kernel.cpp:1723-1736
	f2 n[Q];
#pragma unroll
	for(int i=1; i<Q; i++) n[i] = f[i]-feq[i];
	const f2 Hxx = ((((((((n[1]+n[2])+n[7])+n[8])+n[9])+n[10])+n[13])+n[14])+n[15])+n[16];
	const f2 Hyy = ((((((((n[3]+n[4])+n[7])+n[8])+n[11])+n[12])+n[13])+n[14])+n[17])+n[18];
	const f2 Hzz = ((((((((n[5]+n[6])+n[9])+n[10])+n[11])+n[12])+n[15])+n[16])+n[17])+n[18];
	const f2 Hxy = ((n[7]+n[8])-n[13])-n[14];
	const f2 Hxz = ((n[9]+n[10])-n[15])-n[16];
	const f2 Hyz = ((n[11]+n[12])-n[17])-n[18];
	const float tau0 = __fdiv_rn(1.0f, w0);
	const f2 Qn = ((sm(Hxx, Hxx)+sm(Hyy, Hyy))+sm(Hzz, Hzz))+sm(2.0f, (sm(Hxy, Hxy)+sm(Hxz, Hxz))+sm(Hyz, Hyz));
	return div_rn2(2.0f, bc(tau0)+sqrt_rn2(bc(__fmul_rn(tau0, tau0))+div_rn2(sm(0.76421222f, sqrt_rn2(Qn)), rho)));
}

// f[19] (physical units): streamed DDFs in, post-collision DDFs out
template<uint32_t FEAT> __device__ __forceinline__ void collide_strict2(const DomainConst& c, const StepArgs& a, const PairIn& in, f2* f, PairOut& out) {
	constexpr bool VF = (FEAT&F_VOLUME_FORCE)!=0u, SG = (FEAT&F_SUBGRID)!=0u;
	f2 rho, ux, uy, uz;
	rho_u_strict2(f, rho, ux, uy, uz);
	out.upx = ux; out.upy = uy; out.upz = uz;
	f2 Fin[Q];
	if(VF) {
		const f2 m2rho = sm(-2.0f, rho);
		f2 Fx = bc(a.fx)+sm(m2rho, sm(a.oy, uz)-sm(a.oz, uy));
		f2 Fy = bc(a.fy)+sm(m2rho, sm(a.oz, ux)-sm(a.ox, uz));
		f2 Fz = bc(a.fz)+sm(m2rho, sm(a.ox, uy)-sm(a.oy, ux));
		if(in.zones) zone_force2(in, rho, ux, uy, uz, Fx, Fy, Fz);
		const f2 rho2 = div_rn2(0.5f, rho);
		ux = clampc2(fma2(Fx, rho2, ux)); uy = clampc2(fma2(Fy, rho2, uy)); uz = clampc2(fma2(Fz, rho2, uz));
		forcing_strict2(ux, uy, uz, Fx, Fy, Fz, Fin);
	} else {
		ux = clampc2(ux); uy = clampc2(uy); uz = clampc2(uz);
	}
	out.rho = rho; out.ux = ux; out.uy = uy; out.uz = uz;
	f2 feq[Q];
	feq_strict2(rho, ux, uy, uz, feq);
	f2 w = bc(c.w);
	if(SG) w = smagorinsky_strict2(c.w, f, feq, rho);
	const f2 omw = bc(1.0f)-w;
	if(VF) {
		const f2 c_tau = fma2(w, -0.5f, bc(1.0f));
#pragma unroll
		for(int i=0; i<Q; i++) f[i] = fma2(omw, f[i], fma2(w, feq[i], sm(Fin[i], c_tau)));
	} else {
#pragma unroll
		for(int i=0; i<Q; i++) f[i] = fma2(omw, f[i], fma2(w, feq[i], bc(0.0f)));
	}
}

// =================================================================== FAST: regrouped collision, tolerance-checked
// g[19] = S*f (S = `scale`: 2^15 for FP16S so that the stored half is used as is; 1 otherwise), `inv` = 1/S. In and out in scaled units.
// For each pair of opposite directions (i, i+1), i odd, with a = 3 c_i.u, A = c_i.F, r = w_i*rho:
//   feq_i + feq_i+1 = e = r*(a^2 - 3u^2) + 2 w_i (rho-1)     feq_i - feq_i+1 = 2 r a
//   post-collision: f_i' = (1-w) f_i + U + V,  f_i+1' = (1-w) f_i+1 + U - V,  U = w e/2 + kc (A a/3 + uF),  V = w r a + kc A/3,  kc = 9 w_i (1 - w/2)
// The Smagorinsky tensor H = sum_i c_i c_i (f_i - feq_i) is taken from the second moments of the DDFs instead of from 18 differences:
//   sum c c feq = rho (u u + 1/3)  exactly for this equilibrium, and the DDFs are stored shifted by -w_i (sum c c w = 1/3), hence
//   H_ab = sum c_a c_b f_i  -  rho u_a u_b  -  (rho-1)/3 delta_ab.   Density, momentum and second moments are accumulated in one pass over the pairs.
struct Proj { f2 x, y, z; }; // a vector whose projections on the 9 pair directions are needed
__device__ __forceinline__ f2 proj(const Proj& v, const int k) { // pairs: 0:+x 1:+y 2:+z 3:+x+y 4:+x+z 5:+y+z 6:+x-y 7:+x-z 8:+y-z
	switch(k) {
		case 0: return v.x; case 1: return v.y; case 2: return v.z;
		case 3: return v.x+v.y; case 4: return v.x+v.z; case 5: return v.y+v.z;
		case 6: return v.x-v.y; case 7: return v.x-v.z; default: return v.y-v.z;
	}
}
// HAS_E: some lane of the warp holds a TYPE_E cell (the caller branches warp-uniformly, so that the common path carries no selects)
template<uint32_t FEAT, bool HAS_E> __device__ __forceinline__ void collide_fast2(const DomainConst& c, const StepArgs& a, const PairIn& in, f2* g, const float scale, const float inv, PairOut& out) {
	constexpr bool VF = (FEAT&F_VOLUME_FORCE)!=0u, SG = (FEAT&F_SUBGRID)!=0u;
	f2 rho, rhom1, ir, ux, uy, uz;
	f2 Pxx, Pyy, Pzz, Pxy, Pxz, Pyz; // second moments of the scaled, shifted DDFs (SG only)
	{ // one pass over the pairs: s = f_i + f_i+1 feeds density and second moments, d = f_i - f_i+1 feeds momentum
		f2 R = g[0], mx, my, mz;
#pragma unroll
		for(int k=0; k<9; k++) {
			const f2 sk = g[2*k+1]+g[2*k+2], dk = g[2*k+1]-g[2*k+2];
			R = R+sk;
			switch(k) {
				case 0: mx = dk; if(SG) Pxx = sk; break;
				case 1: my = dk; if(SG) Pyy = sk; break;
				case 2: mz = dk; if(SG) Pzz = sk; break;
				case 3: mx = mx+dk; my = my+dk; if(SG) { Pxx = Pxx+sk; Pyy = Pyy+sk; Pxy = sk; } break;
				case 4: mx = mx+dk; mz = mz+dk; if(SG) { Pxx = Pxx+sk; Pzz = Pzz+sk; Pxz = sk; } break;
				case 5: my = my+dk; mz = mz+dk; if(SG) { Pyy = Pyy+sk; Pzz = Pzz+sk; Pyz = sk; } break;
				case 6: mx = mx+dk; my = my-dk; if(SG) { Pxx = Pxx+sk; Pyy = Pyy+sk; Pxy = Pxy-sk; } break;
				case 7: mx = mx+dk; mz = mz-dk; if(SG) { Pxx = Pxx+sk; Pzz = Pzz+sk; Pxz = Pxz-sk; } break;
				default: my = my+dk; mz = mz-dk; if(SG) { Pyy = Pyy+sk; Pzz = Pzz+sk; Pyz = Pyz-sk; } break;
			}
		}
		rhom1 = inv*R;
		rho = rhom1+bc(1.0f);
		ir = rcp2(rho);
		ir = fma2(ir, fma2(-rho, ir, bc(1.0f)), ir); // one Newton step: full single precision
		const f2 iri = inv*ir;
		ux = mx*iri; uy = my*iri; uz = mz*iri;
	}
	const bool any_e = HAS_E&&(in.e0||in.e1);
	if(any_e) { // TYPE_E lanes: rho/u come from the boundary fields (FX/kernel.cpp:1503-1515)
		rho = sel2(in.e0, in.e1, in.rho_e, rho); ux = sel2(in.e0, in.e1, in.ux_e, ux); uy = sel2(in.e0, in.e1, in.uy_e, uy); uz = sel2(in.e0, in.e1, in.uz_e, uz);
		rhom1 = sel2(in.e0, in.e1, rho-bc(1.0f), rhom1); // the non-TYPE_E partner of a mixed pair keeps its own rho-1: its result must not depend on who it is paired with
		ir = rcp2(rho); ir = fma2(ir, fma2(-rho, ir, bc(1.0f)), ir);
	}
	out.upx = ux; out.upy = uy; out.upz = uz;
	Proj F; f2 uF3 = bc(0.0f);
	if(VF) {
		const f2 m2rho = -2.0f*rho;
		F.x = fma2(m2rho, fma2(a.oy, uz, -(a.oz*uy)), bc(a.fx));
		F.y = fma2(m2rho, fma2(a.oz, ux, -(a.ox*uz)), bc(a.fy));
		F.z = fma2(m2rho, fma2(a.ox, uy, -(a.oy*ux)), bc(a.fz));
		if(in.zones) zone_force2(in, rho, ux, uy, uz, F.x, F.y, F.z);
		const f2 rho2 = 0.5f*ir;
		ux = clampc2(fma2(F.x, rho2, ux)); uy = clampc2(fma2(F.y, rho2, uy)); uz = clampc2(fma2(F.z, rho2, uz));
		uF3 = -(fma2(ux, F.x, fma2(uy, F.y, uz*F.z))); // = 3*uF = -(u.F)
	} else {
		ux = clampc2(ux); uy = clampc2(uy); uz = clampc2(uz);
	}
	out.rho = rho; out.ux = ux; out.uy = uy; out.uz = uz;
	f2 w = bc(c.w);
	if(SG) { // Smagorinsky-Lilly (FX/kernel.cpp:1723-1736) from the second moments
		const f2 rux = rho*ux, ruy = rho*uy, ruz = rho*uz, r3 = 0.33333334f*rhom1;
		const f2 Hxx = fma2(Pxx, inv, -fma2(rux, ux, r3)), Hyy = fma2(Pyy, inv, -fma2(ruy, uy, r3)), Hzz = fma2(Pzz, inv, -fma2(ruz, uz, r3));
		const f2 Hxy = fma2(Pxy, inv, -(rux*uy)), Hxz = fma2(Pxz, inv, -(rux*uz)), Hyz = fma2(Pyz, inv, -(ruy*uz));
		const f2 Qn = fma2(Hxx, Hxx, fma2(Hyy, Hyy, Hzz*Hzz))+2.0f*fma2(Hxy, Hxy, fma2(Hxz, Hxz, Hyz*Hyz));
		const f2 den = bc(c.tau0)+sqrt2(fma2(0.76421222f*sqrt2(Qn), ir, bc(c.tau0sq))); // tau0 = 1/def_w and its square, rounded like the kernel would (host, luw_cabi.cu)
		f2 id = rcp2(den);
		id = fma2(id, fma2(-den, id, bc(1.0f)), id);
		w = 2.0f*id;
	}
	if(any_e) w = sel2(in.e0, in.e1, bc(1.0f), w); // ... and f := feq: relaxation rate 1, no forcing term (FX/kernel.cpp:1747)
	const f2 c3 = -3.0f*fma2(ux, ux, fma2(uy, uy, uz*uz));
	Proj A3; A3.x = 3.0f*ux; A3.y = 3.0f*uy; A3.z = 3.0f*uz; // a = 3 c.u
	const f2 omw = bc(1.0f)-w;
	const f2 hw = (0.5f*scale)*w; // S*w/2
	const f2 hwr = hw*rho; // S*w*rho/2
	const f2 hws = WS*hwr, hwe = WE*hwr; // S*w*r/2 per weight class
	const f2 wrs = 2.0f*hws, wre = 2.0f*hwe; // S*w*r
	const f2 hw1 = hw*rhom1;
	const f2 h1s = (2.0f*WS)*hw1, h1e = (2.0f*WE)*hw1; // S*w/2 * 2 w_i (rho-1)
	const f2 feq0 = W0*fma2(rho, 0.5f*c3, rhom1);
	if(VF) {
		f2 c_tau = fma2(w, -0.5f, bc(1.0f));
		if(any_e) c_tau = sel2(in.e0, in.e1, bc(0.0f), c_tau);
		const f2 kcs = (9.0f*WS/3.0f*scale)*c_tau, kce = (9.0f*WE/3.0f*scale)*c_tau; // S*kc/3
#pragma unroll
		for(int k=0; k<9; k++) {
			const f2 ak = proj(A3, k), Ak = proj(F, k), kc = k<3 ? kcs : kce;
			const f2 U = fma2(kc, fma2(Ak, ak, uF3), fma2(k<3 ? hws : hwe, fma2(ak, ak, c3), k<3 ? h1s : h1e));
			const f2 V = fma2(k<3 ? wrs : wre, ak, kc*Ak);
			const f2 gi = fma2(omw, g[2*k+1], U+V), gj = fma2(omw, g[2*k+2], U-V);
			g[2*k+1] = gi; g[2*k+2] = gj;
		}
		g[0] = fma2(omw, g[0], fma2(2.0f*hw, feq0, ((9.0f*W0/3.0f*scale)*c_tau)*uF3));
	} else {
#pragma unroll
		for(int k=0; k<9; k++) {
			const f2 ak = proj(A3, k);
			const f2 U = fma2(k<3 ? hws : hwe, fma2(ak, ak, c3), k<3 ? h1s : h1e), V = (k<3 ? wrs : wre)*ak;
			const f2 gi = fma2(omw, g[2*k+1], U+V), gj = fma2(omw, g[2*k+2], U-V);
			g[2*k+1] = gi; g[2*k+2] = gj;
		}
		g[0] = fma2(omw, g[0], (2.0f*hw)*feq0);
	}
}

// ------------------------------------------------------------------- FAST, two passes over the DDFs (low register footprint)
// The regrouped collision split so that the 19 DDF pairs never have to be live at once: pass 1 accumulates density, momentum and second moments
// (moments_of), fast_prepare turns them into the per-cell coefficients of the relaxation, pass 2 re-reads the pairs from shared memory and relaxes
// them (fast_relax_axis / fast_relax_diag). TYPE_E lanes are not handled here (the tile kernels overwrite them afterwards).
//
// With G = w rho u + ct F, H = G + ct F (ct = 1 - w/2) the post-collision populations of the pair along c (weight w_i) are
//   f_i'   = (1-w) f_i   + U + V,   f_i+1' = (1-w) f_i+1 + U - V,   U = w_i [ 9/2 (c.u)(c.H) + K0 ],   V = 3 w_i c.G,
//   K0 = w rho (-3/2 u^2) + w (rho-1) - 3 ct u.F        (the same U, V as in collide_fast2, written as a quadratic form in c)
// so that all 18 add terms come from 12 per-cell values: for the axis pairs U = 2 J_a + 2 K0', V = 2 B_a, and for the two diagonal pairs of a plane (a, b)
//   U(a+b) = E_ab + X_ab,  U(a-b) = E_ab - X_ab,  V(a+-b) = B_a +- B_b,   J_a = 9/2 we u_a H_a,  E_ab = J_a + J_b + K0',  X_ab = 9/2 we (u_a H_b + u_b H_a),
//   B_a = 3 we G_a,  K0' = we K0      (we = 1/36 = ws/2 = w0/12; everything in units of the DDF scale S).
struct Moments { f2 R, mx, my, mz, Pxx, Pyy, Pzz, Pxy, Pxz, Pyz; };
// `ld(k, gi, gj)` delivers the decoded pair k (k = 0..8: +x +y +z +x+y +x+z +y+z +x-y +x-z +y-z); the two diagonal pairs of a plane are combined before they are accumulated
template<bool SG, class LD> __device__ __forceinline__ void moments_of(const f2 g0, LD&& ld, Moments& M) {
	f2 gi, gj;
	ld(0, gi, gj); const f2 s0 = gi+gj, d0 = gi-gj;
	ld(1, gi, gj); const f2 s1 = gi+gj, d1 = gi-gj;
	ld(2, gi, gj); const f2 s2 = gi+gj, d2 = gi-gj;
	ld(3, gi, gj); const f2 s3 = gi+gj, d3 = gi-gj;
	ld(6, gi, gj); const f2 s6 = gi+gj, d6 = gi-gj;
	const f2 Sxy = s3+s6, Dxyp = d3+d6, Dxym = d3-d6;
	if(SG) M.Pxy = s3-s6;
	ld(4, gi, gj); const f2 s4 = gi+gj, d4 = gi-gj;
	ld(7, gi, gj); const f2 s7 = gi+gj, d7 = gi-gj;
	const f2 Sxz = s4+s7, Dxzp = d4+d7, Dxzm = d4-d7;
	if(SG) M.Pxz = s4-s7;
	ld(5, gi, gj); const f2 s5 = gi+gj, d5 = gi-gj;
	ld(8, gi, gj); const f2 s8 = gi+gj, d8 = gi-gj;
	const f2 Syz = s5+s8, Dyzp = d5+d8, Dyzm = d5-d8;
	if(SG) M.Pyz = s5-s8;
	M.mx = (d0+Dxyp)+Dxzp; M.my = (d1+Dxym)+Dyzp; M.mz = (d2+Dxzm)+Dyzm;
	M.R = (((g0+s0)+(s1+s2))+(Sxy+Sxz))+Syz;
	if(SG) { M.Pxx = (s0+Sxy)+Sxz; M.Pyy = (s1+Sxy)+Syz; M.Pzz = (s2+Sxz)+Syz; }
}
struct FastK { f2 omw, g0add, UA[3], B[3], E[3], X[3]; }; // see the derivation above; planes: 0 = (x,y), 1 = (x,z), 2 = (y,z)
template<uint32_t FEAT> __device__ __forceinline__ void fast_prepare(const DomainConst& c, const StepArgs& a, const PairIn& in, const Moments& M, const float scale, const float inv, FastK& K, PairOut& out) {
	constexpr bool VF = (FEAT&F_VOLUME_FORCE)!=0u, SG = (FEAT&F_SUBGRID)!=0u;
	constexpr bool EQ = (FEAT&F_EQUILIBRIUM)!=0u;
	// TYPE_E lanes (FX/kernel.cpp:1503-1515, 1747): rho / u are the boundary fields' values, the force half-step and the clamp apply, f := feq -- which is this relaxation
	// with rate 1 and without the forcing term: (1-w) = 0 wipes the streamed-in DDFs, U +- V is the equilibrium. A handful of selects instead of a scalar f_eq per cell.
	const bool any_e = EQ&&in.any_e;
	f2 rhom1 = inv*M.R;
	f2 rho = rhom1+bc(1.0f);
	if(any_e) { // a non-TYPE_E partner keeps its own values: its result must not depend on who it is paired with
		const f2 rho_e = mk2(in.e0 ? c.rho[in.n] : 1.0f, in.e1 ? c.rho[in.n+1ull] : 1.0f);
		rho = sel2(in.e0, in.e1, rho_e, rho); rhom1 = sel2(in.e0, in.e1, rho_e-bc(1.0f), rhom1);
	}
	f2 ir = rcp2(rho);
	ir = fma2(ir, fma2(-rho, ir, bc(1.0f)), ir); // one Newton step: full single precision
	const f2 iri = inv*ir;
	f2 ux = M.mx*iri, uy = M.my*iri, uz = M.mz*iri;
	if(any_e) {
		if(in.e0) { ux.v.x = c.u[in.n]; uy.v.x = c.u[c.N+in.n]; uz.v.x = c.u[2ull*c.N+in.n]; }
		if(in.e1) { ux.v.y = c.u[in.n+1ull]; uy.v.y = c.u[c.N+in.n+1ull]; uz.v.y = c.u[2ull*c.N+in.n+1ull]; }
	}
	out.upx = ux; out.upy = uy; out.upz = uz;
	Proj F; F.x = F.y = F.z = bc(0.0f);
	f2 uF = bc(0.0f);
	if(VF) {
		const f2 m2rho = -2.0f*rho;
		F.x = fma2(m2rho, fma2(a.oy, uz, -(a.oz*uy)), bc(a.fx));
		F.y = fma2(m2rho, fma2(a.oz, ux, -(a.ox*uz)), bc(a.fy));
		F.z = fma2(m2rho, fma2(a.ox, uy, -(a.oy*ux)), bc(a.fz));
		if(in.zones) zone_force2(in, rho, ux, uy, uz, F.x, F.y, F.z);
		const f2 rho2 = 0.5f*ir;
		ux = clampc2(fma2(F.x, rho2, ux)); uy = clampc2(fma2(F.y, rho2, uy)); uz = clampc2(fma2(F.z, rho2, uz));
		uF = fma2(ux, F.x, fma2(uy, F.y, uz*F.z));
	} else {
		ux = clampc2(ux); uy = clampc2(uy); uz = clampc2(uz);
	}
	out.rho = rho; out.ux = ux; out.uy = uy; out.uz = uz;
	f2 w = bc(c.w);
	if(SG) { // Smagorinsky-Lilly (FX/kernel.cpp:1723-1736) from the second moments
		const f2 rux = rho*ux, ruy = rho*uy, ruz = rho*uz, r3 = 0.33333334f*rhom1;
		const f2 Hxx = fma2(M.Pxx, inv, -fma2(rux, ux, r3)), Hyy = fma2(M.Pyy, inv, -fma2(ruy, uy, r3)), Hzz = fma2(M.Pzz, inv, -fma2(ruz, uz, r3));
		const f2 Hxy = fma2(M.Pxy, inv, -(rux*uy)), Hxz = fma2(M.Pxz, inv, -(rux*uz)), Hyz = fma2(M.Pyz, inv, -(ruy*uz));
		const f2 Qn = fma2(Hxx, Hxx, fma2(Hyy, Hyy, Hzz*Hzz))+2.0f*fma2(Hxy, Hxy, fma2(Hxz, Hxz, Hyz*Hyz));
		const f2 den = bc(c.tau0)+sqrt2(fma2(0.76421222f*sqrt2(Qn), ir, bc(c.tau0sq))); // tau0 = 1/def_w and its square, rounded like the kernel would (host, luw_cabi.cu)
		f2 id = rcp2(den);
		id = fma2(id, fma2(-den, id, bc(1.0f)), id);
		w = 2.0f*id;
	}
	if(any_e) w = sel2(in.e0, in.e1, bc(1.0f), w);
	K.omw = bc(1.0f)-w;
	const float swe = scale*WE;
	const f2 wr = w*rho;
	const f2 t1 = fma2(rho, -1.5f*fma2(ux, ux, fma2(uy, uy, uz*uz)), rhom1);
	const f2 hw = swe*w;
	Proj G, H;
	f2 K0;
	if(VF) {
		f2 ct = fma2(w, -0.5f, bc(1.0f));
		if(any_e) ct = sel2(in.e0, in.e1, bc(0.0f), ct);
		const f2 cfx = ct*F.x, cfy = ct*F.y, cfz = ct*F.z;
		G.x = fma2(wr, ux, cfx); G.y = fma2(wr, uy, cfy); G.z = fma2(wr, uz, cfz);
		H.x = G.x+cfx; H.y = G.y+cfy; H.z = G.z+cfz;
		K0 = fma2(hw, t1, ((-3.0f*swe)*ct)*uF);
	} else {
		G.x = wr*ux; G.y = wr*uy; G.z = wr*uz;
		H = G;
		K0 = hw*t1;
	}
	const f2 qx = (4.5f*swe)*ux, qy = (4.5f*swe)*uy, qz = (4.5f*swe)*uz;
	const f2 Jx = qx*H.x, Jy = qy*H.y, Jz = qz*H.z;
	K.X[0] = fma2(qx, H.y, qy*H.x); K.X[1] = fma2(qx, H.z, qz*H.x); K.X[2] = fma2(qy, H.z, qz*H.y);
	K.E[0] = (Jx+Jy)+K0; K.E[1] = (Jx+Jz)+K0; K.E[2] = (Jy+Jz)+K0;
	const f2 K02 = 2.0f*K0;
	K.UA[0] = fma2(2.0f, Jx, K02); K.UA[1] = fma2(2.0f, Jy, K02); K.UA[2] = fma2(2.0f, Jz, K02);
	K.B[0] = (3.0f*swe)*G.x; K.B[1] = (3.0f*swe)*G.y; K.B[2] = (3.0f*swe)*G.z;
	K.g0add = 12.0f*K0;
}
// axis pair ax (= pair index 0..2: +x, +y, +z)
__device__ __forceinline__ void fast_relax_axis(const FastK& K, const int ax, f2& gi, f2& gj) {
	gi = fma2(K.omw, gi, fma2(2.0f, K.B[ax], K.UA[ax])); gj = fma2(K.omw, gj, fma2(-2.0f, K.B[ax], K.UA[ax]));
}
// the two diagonal pairs of plane pl: (gip, gjp) along e_a + e_b (pair 3+pl), (gim, gjm) along e_a - e_b (pair 6+pl); (a, b) = (x,y), (x,z), (y,z)
__device__ __forceinline__ void fast_relax_diag(const FastK& K, const int pl, f2& gip, f2& gjp, f2& gim, f2& gjm) {
	const int ia = pl==2 ? 1 : 0, ib = pl==0 ? 1 : 2;
	const f2 p = K.E[pl]+K.X[pl], q = K.E[pl]-K.X[pl], r = K.B[ia]+K.B[ib], s = K.B[ia]-K.B[ib];
	gip = fma2(K.omw, gip, p+r); gjp = fma2(K.omw, gjp, p-r);
	gim = fma2(K.omw, gim, q+s); gjm = fma2(K.omw, gjm, q-s);
}
// the order in which both passes visit the pairs: axis pairs, then the planes' diagonal pairs together
__host__ __device__ constexpr int fast_pair_order(const int j) { return j<3 ? j : (j-3)%2==0 ? 3+(j-3)/2 : 6+(j-3)/2; } // 0 1 2 3 6 4 7 5 8

} // anonymous namespace
} // namespace luw
