// stream_collide, FAST arithmetic: the "lean loop" variant of the TMA-staged tile kernel (lbm_tile.cuh; re-implements FX/kernel.cpp:1475-1780).
//
// Same data movement, same shared-memory layout, same mbarrier protocol and the same per-cell arithmetic (moments_of / fast_prepare / fast_relax_*, lbm_vec.cuh)
// as k_stream_collide_tile<.., FAST = true> with a two-pass configuration -- results are bit-identical to it. What differs is how a consumer warp spends its
// issue slots. The round-1 kernel was issue-bound on the urban LES step (930 warp instructions per 64 cells, of which ~200 were per-tile control flow,
// address arithmetic, flag decoding and store masks, ~95 relaxation-zone gathers and ~125 producer work and polling; profiles/r1c_ncu_urban_fp16s.md).
// Here (profiles/r2a_ncu_urban_fp16s_v5.md is the first cut of this kernel, profiles/r2b_* the current one):
//   * everything that depends only on the strip (row coordinates, halo / boundary classification, the y/z part of the relaxation zones) is computed when a
//     strip starts, not per tile; shared memory is addressed through 32-bit shared-window addresses with immediate box offsets;
//   * per tile a warp takes ONE warp-uniform decision: are all 64 cells plain fluid or TYPE_E (flag byte 0x00 / 0x02), in a tile without halo columns or
//     columns beyond the lattice?  Then it runs the fast body: no run masks -- pass 2 stores whole words unconditionally; relaxation-zone data is gathered
//     only by warps that reach into a zone (per-row part once, per-cell part for west / east), TYPE_E lanes are overwritten afterwards by the out-of-line
//     equilibrium. Only warps that hold solid / gas cells, halo columns or a partial tile go through the general body (out of line), which is the round-1
//     logic with the same arithmetic, so a cell's result does not depend on the path its warp takes;
//   * the periodic-x column is parked by the thread that reads it back at the end of the strip (the row's last pair), so the consumers of a CTA never
//     rendezvous (the round-1 kernel has a CTA-wide named barrier per strip for this: 5 % of all stall samples);
//   * only the warp that holds a row's last pair waits for the NEXT stage before it collides (it reads column 0 of the next tile);
//   * the producer warp blocks in mbarrier.try_wait (which suspends in hardware) without a software back-off or spin counter around it: a lost arrival is
//     caught by the consumers' bounded waits.
#pragma once
#include "lbm_tile.cuh"

namespace luw {
namespace {

// ------------------------------------------------------------------ relaxation zones, split by what depends on x (FX/kernel.cpp:1523-1614)
struct ZoneRow { // of one lattice row (y, z): the part of zone_prefetch that does not depend on x
	uint32_t d_yz; // distance to the nearest of the south / north / top faces in whose nudging shell the row lies (first minimum in the reference's order s, n, t); buffer_N+1 if none
	uint64_t ref_yz; // that face's reference cell for x = 0
	bool sponge; float ks; uint64_t ref_top; // top sponge: the row lies in it, its sigma, the top-plane reference cell for x = 0
};
__device__ __forceinline__ ZoneRow zone_row(const DomainConst& c, const uint32_t y, const uint32_t z) {
	ZoneRow r;
	const uint64_t row = c.Px, plane = (uint64_t)c.Px*c.Ny;
	r.d_yz = c.buffer_N+1u; r.ref_yz = 0ull; r.sponge = false; r.ks = 0.0f; r.ref_top = (uint64_t)y*row+(uint64_t)c.tz*plane;
	if(c.features&F_NUDGING) {
		const int yg = (int)y+c.Oy, zg = (int)z+c.Oz, Nb = (int)c.buffer_N;
		const int ds = yg, dn = (int)(c.Nyg-1u)-yg, dt = (int)(c.Nzg-1u)-zg;
		if(c.downstream_face!=3&&c.has_s&&ds>=0&&ds<=Nb&&(uint32_t)ds<r.d_yz) { r.d_yz = (uint32_t)ds; r.ref_yz = (uint64_t)c.sy*row+(uint64_t)z*plane; }
		if(c.downstream_face!=4&&c.has_n&&dn>=0&&dn<=Nb&&(uint32_t)dn<r.d_yz) { r.d_yz = (uint32_t)dn; r.ref_yz = (uint64_t)c.ny*row+(uint64_t)z*plane; }
		if(c.has_t&&dt>=0&&dt<=Nb&&(uint32_t)dt<r.d_yz) { r.d_yz = (uint32_t)dt; r.ref_yz = r.ref_top; }
	}
	if((c.features&F_SPONGE)&&c.has_t) {
		const int dt = (int)(c.Nzg-2u)-((int)z+c.Oz);
		if(dt>=0&&dt<(int)c.sponge_N) { r.sponge = true; r.ks = __ldg(c.sigma+dt); }
	}
	return r;
}
// the per-cell part: west / east shells, then the loads. Produces exactly zone_prefetch's values (first minimum in the order w, e, s, n, t).
__device__ __forceinline__ ZoneRef zone_cell(const DomainConst& c, const ZoneRow& zr, const uint32_t x, const uint64_t n_row, const bool active) {
	ZoneRef r;
	r.nudge = false; r.sponge = false; r.kn = 0.0f; r.unx = r.uny = r.unz = 0.0f; r.ks = 0.0f; r.usx = r.usy = r.usz = 0.0f;
	if(!active) return r;
	if(c.features&F_NUDGING) {
		const int xg = (int)x+c.Ox, Nb = (int)c.buffer_N;
		const int dw = xg, de = (int)(c.Nxg-1u)-xg;
		uint32_t dmin = c.buffer_N+1u;
		uint64_t nref = 0ull;
		if(c.downstream_face!=1&&c.has_w&&dw>=0&&dw<=Nb) { dmin = (uint32_t)dw; nref = (uint64_t)c.wx+n_row; }
		if(c.downstream_face!=2&&c.has_e&&de>=0&&de<=Nb&&(uint32_t)de<dmin) { dmin = (uint32_t)de; nref = (uint64_t)c.ex+n_row; }
		if(zr.d_yz<dmin) { dmin = zr.d_yz; nref = (uint64_t)x+zr.ref_yz; }
		if(dmin<=c.buffer_N) {
			r.nudge = true;
			r.kn = __fmul_rn(__ldg(c.wbuf+dmin), c.buffer_inv_tau);
			r.unx = __ldg(c.u+nref); r.uny = __ldg(c.u+c.N+nref); r.unz = __ldg(c.u+2ull*c.N+nref);
		}
	}
	if(zr.sponge) {
		const uint64_t nref = (uint64_t)x+zr.ref_top;
		r.sponge = true;
		r.ks = zr.ks;
		r.usx = __ldg(c.u+nref); r.usy = __ldg(c.u+c.N+nref); r.usz = __ldg(c.u+2ull*c.N+nref);
	}
	return r;
}

// does the row (y, z) lie in a relaxation zone through its y / z position? (the per-strip test of the lean loop: zone_row's two conditions without its loads)
__device__ __forceinline__ bool zone_row_hit(const DomainConst& c, const uint32_t y, const uint32_t z) {
	const int yg = (int)y+c.Oy, zg = (int)z+c.Oz, Nb = (int)c.buffer_N;
	bool hit = false;
	if(c.features&F_NUDGING) {
		const int ds = yg, dn = (int)(c.Nyg-1u)-yg, dt = (int)(c.Nzg-1u)-zg;
		hit = (c.downstream_face!=3&&c.has_s&&ds>=0&&ds<=Nb)||(c.downstream_face!=4&&c.has_n&&dn>=0&&dn<=Nb)||(c.has_t&&dt>=0&&dt<=Nb);
	}
	if((c.features&F_SPONGE)&&c.has_t) { const int dt = (int)(c.Nzg-2u)-zg; hit = hit||(dt>=0&&dt<(int)c.sponge_N); }
	return hit;
}

// ------------------------------------------------------------------ shared memory by 32-bit shared-window address
// (ptxas folds `address + constant` into the instruction's immediate offset, so one base register serves all boxes of a stage)
__device__ __forceinline__ uint32_t lds_b32(const uint32_t a) { uint32_t v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u16(const uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ float lds_f32(const uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ float2 lds_f32x2(const uint32_t a) { float2 v; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ void sts_b32(const uint32_t a, const uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_b16(const uint32_t a, const uint32_t v) { asm volatile("st.shared.b16 [%0], %1;" :: "r"(a), "r"(v) : "memory"); } // the low half of v
__device__ __forceinline__ void sts_f32(const uint32_t a, const float v) { asm volatile("st.shared.f32 [%0], %1;" :: "r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_f32x2(const uint32_t a, const float2 v) { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" :: "r"(a), "f"(v.x), "f"(v.y) : "memory"); }
__device__ __forceinline__ bool mbar_try_a(const uint32_t bar, const uint32_t parity) {
	uint32_t done;
	asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity), "r"(1000000u) : "memory");
	return done!=0u;
}
__device__ __forceinline__ void mbar_wait_a(const uint32_t bar, const uint32_t parity) {
	if(mbar_try_a(bar, parity)) return;
	uint32_t spins = 0u;
	while(!mbar_try_a(bar, parity)) if(++spins>(1u<<24)) __trap(); // a lost arrival must abort the launch, not hang the device
}
// ... yielding: a consumer that has run ahead of its CTA to the end of the ring polls while the warps it waits for need the issue slots (experiment: LC_YIELD)
__device__ __forceinline__ void mbar_wait_y(const uint32_t bar, const uint32_t parity, const bool yield) {
	if(mbar_try_a(bar, parity)) return;
	uint32_t spins = 0u;
	while(!mbar_try_a(bar, parity)) { if(yield) __nanosleep(128u); if(++spins>(1u<<24)) __trap(); }
}
__device__ __forceinline__ void mbar_arrive_a(const uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }

// the pair's word of a box (R), one element of it (E), and the FAST codecs (FP16S: the stored half IS the scaled value)
template<int P> struct SmemPair;
template<> struct SmemPair<P_FP32> {
	typedef float2 R; typedef float E;
	static __device__ __forceinline__ R ldw(const uint32_t a) { return lds_f32x2(a); }
	static __device__ __forceinline__ E lde(const uint32_t a) { return lds_f32(a); }
	static __device__ __forceinline__ void stw(const uint32_t a, const R w) { sts_f32x2(a, w); }
	static __device__ __forceinline__ void ste(const uint32_t a, const E e) { sts_f32(a, e); }
	static __device__ __forceinline__ E low(const R w) { return w.x; }
	static __device__ __forceinline__ R shift_in(const R w0, const E next) { return make_float2(w0.y, next); }
	static __device__ __forceinline__ R join(const E e0, const E e1) { return make_float2(e0, e1); }
	static __device__ __forceinline__ void shift_out_both(const uint32_t own, const uint32_t next, const R n) { sts_f32(own+4u, n.x); sts_f32(next, n.y); }
	static __device__ __forceinline__ f2 dec(const R w) { f2 v; v.v = w; return v; }
	static __device__ __forceinline__ R enc(const f2 v) { return v.v; }
};
template<> struct SmemPair<P_FP16S> {
	typedef uint32_t R; typedef uint32_t E; // an element travels in the low half of a 32-bit register
	static __device__ __forceinline__ R ldw(const uint32_t a) { return lds_b32(a); }
	static __device__ __forceinline__ E lde(const uint32_t a) { return lds_u16(a); }
	static __device__ __forceinline__ void stw(const uint32_t a, const R w) { sts_b32(a, w); }
	static __device__ __forceinline__ void ste(const uint32_t a, const E e) { sts_b16(a, e); }
	static __device__ __forceinline__ E low(const R w) { return w; }
	static __device__ __forceinline__ R shift_in(const R w0, const E next) { return __byte_perm(w0, next, 0x5432); }
	static __device__ __forceinline__ R join(const E e0, const E e1) { return __byte_perm(e0, e1, 0x5410); }
	static __device__ __forceinline__ void shift_out_both(const uint32_t own, const uint32_t next, const R n) { sts_b16(own+2u, n); sts_b16(next, n>>16); }
	static __device__ __forceinline__ f2 dec(const R w) { return PairCodec<P_FP16S>::dec_raw(w); }
	static __device__ __forceinline__ R enc(const f2 v) { return PairCodec<P_FP16S>::enc_raw(v); }
};
template<> struct SmemPair<P_FP16C> : SmemPair<P_FP16S> {
	static __device__ __forceinline__ f2 dec(const R w) { return PairCodec<P_FP16C>::dec(w); }
	static __device__ __forceinline__ R enc(const f2 v) { return PairCodec<P_FP16C>::enc_fast(v); }
};

// rho/u of TYPE_E cells: into L2 one tile ahead (`fn`: the pair's flags in the NEXT stage)
template<class CFG> __device__ __noinline__ void lean_prefetch_e(const DomainConst& c, const uint32_t fn, const uint64_t n) {
	if((fn&0x0003u)==TYPE_E||(fn&0x0300u)==(TYPE_E<<8)) { const uint64_t m = n+(uint64_t)CFG::TX; prefetch_l2(c.rho+m); prefetch_l2(c.u+m); prefetch_l2(c.u+c.N+m); prefetch_l2(c.u+2ull*c.N+m); }
}

// The general body of the lean loop: run masks (solid / gas cells, halo columns, columns beyond the lattice), TYPE_E cells, relaxation zones. Out of line: about a
// quarter of the warp-tiles of an urban case come here, and keeping it out of the loop keeps the loop's registers and instruction-cache footprint for the fast body.
// Same arithmetic, in the same order, as the fast body (a cell's result must not depend on the path its warp takes). bb / nxt: shared-window addresses.
template<class CFG, uint32_t FEAT> __device__ __noinline__ void lean_general_pair(const DomainConst& c, const StepArgs& a, const uint32_t bb, const uint32_t nxt, const uint32_t fl2,
	const uint32_t x, const uint32_t y, const uint32_t z, const bool zone_warp, const bool odd_end) {
	constexpr int P = CFG::P;
	typedef SmemPair<P> SP;
	typedef typename SP::R R;
	typedef PairCodec<P> PC;
	constexpr bool UF = (FEAT&F_UPDATE_FIELDS)!=0u, EQ = (FEAT&F_EQUILIBRIUM)!=0u, SG = (FEAT&F_SUBGRID)!=0u, TH = (FEAT&F_TEMPERATURE)!=0u;
	const float scale = (P==P_FP16S) ? 32768.0f : 1.0f, inv = (P==P_FP16S) ? 3.0517578E-5f : 1.0f;
	const uint32_t fl0 = fl2&0xFFu, fl1 = fl2>>8;
	bool run0 = !((fl0&TYPE_BO)==TYPE_S||(fl0&TYPE_SU)==TYPE_G), run1 = !((fl1&TYPE_BO)==TYPE_S||(fl1&TYPE_SU)==TYPE_G);
	run0 = run0&&x<c.Nx&&!(c.Dx>1u&&(x==0u||x>=c.Nx-1u));
	run1 = run1&&x+1u<c.Nx&&!(c.Dx>1u&&(x+1u>=c.Nx-1u));
	if(!(run0||run1)) return;
	const uint64_t n_row = (uint64_t)y*c.Px+(uint64_t)z*((uint64_t)c.Px*c.Ny), n = n_row+(uint64_t)x;
	const uint32_t bo0 = fl0&TYPE_BO, bo1 = fl1&TYPE_BO;
	const bool e0 = EQ&&run0&&bo0==TYPE_E, e1 = EQ&&run1&&bo1==TYPE_E;
	PairIn in;
	in.zones = false;
	in.e0 = e0; in.e1 = e1; in.any_e = e0||e1; in.n = n; // TYPE_E lanes: rho / u are the boundary fields' (loaded in fast_prepare; prefetched into L2 a tile ago)
	if(zone_warp) {
		const ZoneRow zr = zone_row(c, y, z);
		in.nudge_vertical = c.nudge_vertical;
		in.zr0 = zone_cell(c, zr, x, n_row, run0&&bo0!=TYPE_E);
		in.zr1 = zone_cell(c, zr, x+1u, n_row, run1&&bo1!=TYPE_E);
		in.zones = in.zr0.nudge||in.zr0.sponge||in.zr1.nudge||in.zr1.sponge;
	}
	Moments M;
	const f2 g0 = SP::dec(SP::ldw(bb));
	const uint32_t sh0 = odd_end ? nxt : bb+(uint32_t)CFG::ES; // cell 0's element of an x-shifted box (+ box_off): the second element of the pair's word -- or, odd Nx, the row's last pair: `nxt` (cell 0 IS the last column, cell 1 does not exist)
	const auto ldb = [&](const int k) -> R { // box B: the pair's word, or -- x-shifted -- the second element of its word and the element to the right of it
		if(pair_shifted(k)) return SP::join(SP::lde(sh0+(uint32_t)CFG::box_off(2+2*k)), SP::lde(nxt+(uint32_t)CFG::box_off(2+2*k)));
		return SP::ldw(bb+(uint32_t)CFG::box_off(2+2*k));
	};
	const auto ld1 = [&](const int k, f2& gi, f2& gj) { gi = SP::dec(SP::ldw(bb+(uint32_t)CFG::box_off(1+2*k))); gj = SP::dec(ldb(k)); };
	moments_of<SG>(g0, ld1, M);
	FastK K;
	PairOut out;
	fast_prepare<FEAT>(c, a, in, M, scale, inv, K, out);
	if(TH) store_upre(c, n, run0&&!e0, run1&&!e1, out);
	if(UF) { // rho / u of the non-TYPE_E cells (UPDATE_FIELDS, FX/kernel.cpp:1709-1715)
		const bool w0 = run0&&!e0, w1 = run1&&!e1;
		if(w0&&w1) {
			*(float2*)(c.rho+n) = out.rho.v; *(float2*)(c.u+n) = out.ux.v; *(float2*)(c.u+c.N+n) = out.uy.v; *(float2*)(c.u+2ull*c.N+n) = out.uz.v;
		} else {
			if(w0) { c.rho[n] = out.rho.v.x; c.u[n] = out.ux.v.x; c.u[c.N+n] = out.uy.v.x; c.u[2ull*c.N+n] = out.uz.v.x; }
			if(w1) { c.rho[n+1ull] = out.rho.v.y; c.u[n+1ull] = out.ux.v.y; c.u[c.N+n+1ull] = out.uy.v.y; c.u[2ull*c.N+n+1ull] = out.uz.v.y; }
		}
	}
	const auto mix = [&](const R nw, const R old) -> R { return PC::mix(run0, run1, nw, old); };
	SP::stw(bb, mix(SP::enc(fma2(K.omw, g0, K.g0add)), SP::ldw(bb)));
	struct Raw { R wa, wb0; };
	const auto ld2 = [&](const int k, Raw& r, f2& gi, f2& gj) {
		r.wa = SP::ldw(bb+(uint32_t)CFG::box_off(1+2*k)); r.wb0 = ldb(k);
		gi = SP::dec(r.wa); gj = SP::dec(r.wb0);
	};
	const auto st2 = [&](const int k, const Raw& r, const f2 gi, const f2 gj) { // f_i' goes to slot B, f_i+1' to slot A
		const int bA = 1+2*k, bB = 2+2*k;
		const R ni = SP::enc(gi), nj = SP::enc(gj);
		SP::stw(bb+(uint32_t)CFG::box_off(bA), mix(nj, r.wa));
		if(pair_shifted(k)) {
			if(P==P_FP32) { if(run0) sts_f32(sh0+(uint32_t)CFG::box_off(bB), SmemPair<P_FP32>::enc(gi).x); if(run1) sts_f32(nxt+(uint32_t)CFG::box_off(bB), SmemPair<P_FP32>::enc(gi).y); }
			else { const uint32_t n16 = *(const uint32_t*)&ni; if(run0) sts_b16(sh0+(uint32_t)CFG::box_off(bB), n16); if(run1) sts_b16(nxt+(uint32_t)CFG::box_off(bB), n16>>16); }
		} else SP::stw(bb+(uint32_t)CFG::box_off(bB), mix(ni, r.wb0));
	};
#pragma unroll
	for(int ax=0; ax<3; ax++) {
		Raw r; f2 gi, gj;
		ld2(ax, r, gi, gj);
		fast_relax_axis(K, ax, gi, gj);
		st2(ax, r, gi, gj);
	}
#pragma unroll
	for(int pl=0; pl<3; pl++) {
		Raw rp, rm; f2 gip, gjp, gim, gjm;
		ld2(3+pl, rp, gip, gjp); ld2(6+pl, rm, gim, gjm);
		fast_relax_diag(K, pl, gip, gjp, gim, gjm);
		st2(3+pl, rp, gip, gjp); st2(6+pl, rm, gim, gjm);
	}
}

// ------------------------------------------------------------------ the kernel
// Per-launch constants of the lean loop, derived on the host (lean_const): they reach the kernel through the constant bank, so the loop spends neither
// registers nor instructions on them (the first cut kept them in registers and spilled four of them: 5 % of the stall samples were loads of those spills).
struct LeanConst {
	uint32_t tiles_x, tiles_y, tiles_z;
	uint32_t last_tx; // cells of the last tile's rows that lie inside the lattice
	uint32_t rowend_last; // local x of the pair that holds the last cell of a row, in the strip's last tile
	int zone_xw, zone_xe; // a warp (64 x-consecutive cells from local x = xw) reaches the west nudging shell iff xw <= zone_xw, the east one iff xw >= zone_xe
	int slow_xlo, slow_xhi; // ... holds a halo column or columns beyond the lattice iff xw <= slow_xlo or xw >= slow_xhi: only THESE warps of a strip's first / last tile take the general body
	uint32_t flags;
};
enum : uint32_t { LC_WRAP_X = 1u, LC_PARK = 2u, LC_EDGE_X_SLOW = 4u, LC_ZONES = 8u, LC_PREFETCH = 16u, LC_UF = 32u, LC_EQ = 64u, LC_LAG = 128u, LC_ODD_X = 256u, LC_YIELD = 512u };
template<class CFG> inline LeanConst lean_const(const DomainConst& c, const bool prefetch, const bool lag, const bool yield = false) {
	LeanConst l;
	l.tiles_x = (c.Nx+CFG::TX-1u)/CFG::TX; l.tiles_y = (c.Ny+CFG::TY-1u)/CFG::TY; l.tiles_z = (c.Nz+CFG::TZ-1u)/CFG::TZ;
	l.last_tx = c.Nx-(l.tiles_x-1u)*(uint32_t)CFG::TX;
	l.rowend_last = (l.last_tx-1u)&~1u;
	const bool vf = (c.features&F_VOLUME_FORCE)!=0u, zones = vf&&(c.features&(F_NUDGING|F_SPONGE))!=0u;
	const bool west = zones&&(c.features&F_NUDGING)&&c.downstream_face!=1&&c.has_w, east = zones&&(c.features&F_NUDGING)&&c.downstream_face!=2&&c.has_e;
	l.slow_xlo = c.Dx>1u ? 0 : -1; // x = 0 is a halo column
	l.slow_xhi = (int)(c.Dx>1u ? c.Nx-1u : c.Nx)-63; // the warp's 64 cells end beyond the last executing column
	l.zone_xw = west ? (int)c.buffer_N-c.Ox : -0x7FFFFFFF;
	l.zone_xe = east ? (int)c.Nxg-1-(int)c.buffer_N-63-c.Ox : 0x7FFFFFFF;
	l.flags = (c.Dx==1u ? LC_WRAP_X : 0u)|((c.Dx==1u&&l.tiles_x>=2u) ? LC_PARK : 0u)|((c.Dx>1u||l.last_tx!=(uint32_t)CFG::TX) ? LC_EDGE_X_SLOW : 0u)|(zones ? LC_ZONES : 0u)|(prefetch ? LC_PREFETCH : 0u)
		|((c.features&F_UPDATE_FIELDS) ? LC_UF : 0u)|((c.features&F_EQUILIBRIUM) ? LC_EQ : 0u)|((lag&&CFG::STAGES>=5) ? LC_LAG : 0u)|((c.Nx&1u) ? LC_ODD_X : 0u)|(yield ? LC_YIELD : 0u);
	return l;
}
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, const int c0, const int c1, const int c2, const int c3) {
	asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" :: "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

template<class CFG, uint32_t FEAT> __global__ void __maxnreg__(tile_max_regs(CFG::THREADS/32, CFG::CTAS_PER_SM))
k_stream_collide_lean(const __grid_constant__ DomainConst c, const __grid_constant__ StepArgs a, const __grid_constant__ TileMaps maps, const __grid_constant__ LeanConst lc) {
	const uint32_t tiles_x = lc.tiles_x, tiles_y = lc.tiles_y, tiles_z = lc.tiles_z;
	constexpr int P = CFG::P, TX = CFG::TX, TY = CFG::TY, TZ = CFG::TZ, S = CFG::STAGES, NC = CFG::CONSUMERS;
	typedef PairCodec<P> PC;
	typedef typename PC::R R;
	typedef typename PC::E E;
	constexpr bool UF = (FEAT&F_UPDATE_FIELDS)!=0u, EQ = (FEAT&F_EQUILIBRIUM)!=0u, VF = (FEAT&F_VOLUME_FORCE)!=0u, SG = (FEAT&F_SUBGRID)!=0u, TH = (FEAT&F_TEMPERATURE)!=0u;
	static_assert(TX/2>=32, "a consumer warp must lie inside one lattice row");

	extern __shared__ uint8_t smem_raw[];
	uint8_t* const stage0 = smem_raw+((128u-(smem_u32(smem_raw)&127u))&127u);
	uint64_t* const bar_full = (uint64_t*)(stage0+(size_t)S*CFG::STAGE_BYTES);
	uint64_t* const bar_done = bar_full+S;
	uint64_t* const bar_head = bar_done+S;
	volatile uint32_t* const tile_strip = (volatile uint32_t*)(bar_head+1);
	volatile int* const tile_yz = (volatile int*)(tile_strip+S);

	const uint32_t tid = threadIdx.x;
	if(tid==0u) {
		if(blockIdx.x==0u) c.sched[(uint32_t)(a.t&1ull)^1u] = 0u; // strip counter of the NEXT step
		for(int s=0; s<S; s++) { mbar_init(bar_full+s, 1u); mbar_init(bar_done+s, (uint32_t)(NC/32)); }
		mbar_init(bar_head, 1u);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	const uint32_t nstrips = tiles_y*tiles_z;
	constexpr uint32_t END = 0xFFFFFFFFu, STRIP_BND = 0x80000000u; // tile_strip[s]: strip id | STRIP_BND for a boundary strip of an overlapped halo exchange (DomainConst::so_nb)
	const uint32_t odd = (uint32_t)(a.t&1ull);
	const bool wrap_x = (lc.flags&LC_WRAP_X)!=0u;
	const bool park = (lc.flags&LC_PARK)!=0u;

	if(tid>=(uint32_t)NC) { // ---------------------------------------------------------------- producer warp (protocol of k_stream_collide_tile)
		const bool leader = (tid&31u)==0u;
		uint32_t lstrip = 0u, lxt = 0u, issued = 0u, lbnd = 0u;
		bool ended = false;
		const bool lag = (lc.flags&LC_LAG)!=0u;
		const auto issue_loads = [&]() {
			const int s = (int)(issued%(uint32_t)S);
			if(lxt==0u) {
				uint32_t v = 0u;
				if(leader) v = atomicAdd(c.sched+odd, 1u);
				lstrip = __shfl_sync(0xFFFFFFFFu, v, 0);
				if(lstrip>=nstrips) {
					ended = true;
					if(leader) { tile_strip[s] = END; mbar_arrive(bar_full+s); }
					return;
				}
				lbnd = lstrip<c.so_nb ? STRIP_BND : 0u; // boundary strips come first (strip_of) and are counted in *c.bdone when they are in global memory
				lstrip = strip_of(c, lstrip, tiles_y, tiles_z);
			}
			const int x0 = (int)lxt*TX, y0 = (int)(lstrip%tiles_y)*TY, z0 = (int)(lstrip/tiles_y)*TZ;
			uint8_t* st = stage0+(size_t)s*CFG::STAGE_BYTES;
			if(EQ&&lxt==0u) prefetch_west_face<CFG>(c, tid&31u, y0, z0);
			if(leader) {
				tile_strip[s] = lstrip|lbnd;
				tile_yz[2*s] = y0; tile_yz[2*s+1] = z0;
				mbar_expect_tx(bar_full+s, (uint32_t)CFG::LOAD_BYTES);
				tma_load_3d(st+CFG::FLAG_OFF, &maps.flags, bar_full+s, x0, y0, z0);
				tma_load_4d(st, &maps.fi, bar_full+s, x0, y0, z0, 0);
				tma_load_4d(st+CFG::box_off(1), &maps.fiA, bar_full+s, x0, y0, z0, odd ? 1 : 2);
#pragma unroll
				for(int k=0; k<9; k++) {
					int cx, cy, cz; pair_shift(k, cx, cy, cz);
					const int i = 2*k+1;
					tma_load_4d(st+CFG::box_off(2+2*k), &maps.fi, bar_full+s, x0, y0+cy, z0+cz, odd ? i+1 : i);
				}
				if((lc.flags&LC_PREFETCH)&&lxt+1u<tiles_x) { // the strip's next tile into L2: its stage is still busy, so its loads will be issued a tile-time from now and then hit L2
					const int x1 = x0+TX;
					tma_prefetch_4d(&maps.fi, x1, y0, z0, 0);
					tma_prefetch_4d(&maps.fiA, x1, y0, z0, odd ? 1 : 2);
#pragma unroll
					for(int k=0; k<9; k++) {
						int cx, cy, cz; pair_shift(k, cx, cy, cz);
						const int i = 2*k+1;
						tma_prefetch_4d(&maps.fi, x1, y0+cy, z0+cz, odd ? i+1 : i);
					}
				}
			}
			if(++lxt==tiles_x) lxt = 0u;
			issued++;
		};
		for(int i=0; i<S&&!ended; i++) issue_loads();
		uint32_t sxt = 0u;
		for(uint32_t q=0u; q<issued; q++) {
			const int s = (int)(q%(uint32_t)S);
			{ // try_wait suspends the warp in hardware until the barrier sees traffic or the time hint expires (measured: __nanosleep around it returns at once and only
				// costs issue slots). No spin counter here: if an arrival were lost, the consumers' bounded waits on `full` abort the launch.
				const uint32_t par = (q/(uint32_t)S)&1u;
				while(!mbar_try(bar_done+s, par)) { }
			}
			const int x0 = (int)sxt*TX, y0 = tile_yz[2*s], z0 = tile_yz[2*s+1];
			const uint8_t* st = stage0+(size_t)s*CFG::STAGE_BYTES;
			const bool inner = y0>0&&z0>0&&y0+TY<=(int)c.Ny&&z0+TZ<=(int)c.Nz;
			const bool last_of_strip = sxt+1u==tiles_x;
			if(leader) {
				tma_store_4d(&maps.fi, st, x0, y0, z0, 0);
				tma_store_4d(&maps.fiA, st+CFG::box_off(1), x0, y0, z0, odd ? 1 : 2);
				if(inner) {
#pragma unroll
					for(int k=0; k<9; k++) {
						int cx, cy, cz; pair_shift(k, cx, cy, cz);
						const int i = 2*k+1;
						tma_store_4d(&maps.fi, st+CFG::box_off(2+2*k), x0, y0+cy, z0+cz, odd ? i+1 : i);
					}
				} else {
#pragma unroll
					for(int k=0; k<9; k++) {
						int cx, cy, cz; pair_shift(k, cx, cy, cz);
						const int i = 2*k+1;
						if(!box_by_threads<CFG>(c, cy, cz, y0, z0)) tma_store_4d(&maps.fi, st+CFG::box_off(2+2*k), x0, y0+cy, z0+cz, odd ? i+1 : i);
					}
				}
				tma_commit();
				if(last_of_strip&&(tile_strip[s]&STRIP_BND)!=0u) { tma_wait_all0(); __threadfence(); atomicAdd(c.bdone, 1u); } // a boundary strip is in global memory: the halo exchange may read it
				// refill: the stage just stored may be overwritten once TMA has read it. LC_LAG: do not wait for that here -- refill the stage of the PREVIOUS tile instead, whose
				// stores were committed a tile-time ago (the wait then returns at once; the ring is one tile shallower)
				if(!ended) { if(lag) { if(q>0u) tma_wait_read1(); } else tma_wait_read0(); }
			}
			__syncwarp();
			if(!ended&&(!lag||q>0u)) issue_loads();
			if(park&&last_of_strip&&leader) { tma_wait_all_but(tiles_x-1u); mbar_arrive(bar_head); }
			sxt = last_of_strip ? 0u : sxt+1u;
		}
		if(leader) tma_wait_all0();
		return;
	}

	// ---------------------------------------------------------------------------------------- consumers: two cells per thread
	// Shared memory is addressed through 32-bit shared-window addresses and explicit ld/st.shared (SmemPair): one base register per tile, every box at an
	// immediate offset, no generic-pointer arithmetic in the loop.
	typedef SmemPair<P> SP;
	const uint32_t sm0 = smem_u32(stage0), bar0 = sm0+(uint32_t)(S*CFG::STAGE_BYTES); // full[s] at bar0 + 8 s, done[s] at bar0 + 8 (S + s), head at bar0 + 16 S
	const uint32_t row = tid/(uint32_t)(TX/2), lx = 2u*(tid%(uint32_t)(TX/2)), ly = row%(uint32_t)TY, lz = row/(uint32_t)TY;
	const float scale = (P==P_FP16S) ? 32768.0f : 1.0f, inv = (P==P_FP16S) ? 3.0517578E-5f : 1.0f;
	const bool has_zones = VF&&(lc.flags&LC_ZONES)!=0u;
	const bool next_warp = (lx&~63u)==(uint32_t)(TX-64); // this warp holds the last pair of its row in a full tile: it reads column 0 of the NEXT tile
	constexpr uint32_t E_BITS = TYPE_E|(TYPE_E<<8), T_BITS = TYPE_T|(TYPE_T<<8); // TYPE_T (temperature boundary) means nothing to the momentum step: a thermal deck flags every TYPE_E cell TYPE_T as well

	// per x tile of a strip (bit min(xt, 31)): does this warp hold a halo column / columns beyond the lattice (-> general body), does it reach the west / east nudging shell?
	// Decided once per launch instead of per tile (four constant-bank loads and compares per tile otherwise); past 32 tiles bit 31 answers conservatively.
	const bool end_full = lx==(uint32_t)(TX-2), end_last = lx==lc.rowend_last; // this thread holds the last pair of its row in a full tile / in the strip's last tile
	uint32_t edge_bits = 0u, zonex_bits = 0u;
	for(uint32_t t=0u; t<tiles_x; t++) {
		const int xw = (int)(t*(uint32_t)TX+(lx&~63u));
		const uint32_t bit = 1u<<(t<31u ? t : 31u);
		if(xw<=lc.slow_xlo||xw>=lc.slow_xhi) edge_bits |= bit;
		if(xw<=lc.zone_xw||xw>=lc.zone_xe) zonex_bits |= bit;
	}
	uint32_t s = 0u, ph = 0u, kstrip = 0u, st = sm0, xt = 0u; // ring slot, its phase, strips done, shared address of stage s, x tile inside the strip
	int y0 = 0, z0 = 0, py0 = 0, pz0 = 0;
	uint32_t y = 0u, z = 0u, park_off = 0u;
	uint32_t sf = 0u; // strip state in one register: the strip touches the y/z boundary (uniform in the CTA) | the cells of this row execute | the row lies in a relaxation zone through its y / z position (both uniform in the warp)
	constexpr uint32_t SF_BND = 1u, SF_IN = 2u, SF_ZONE = 4u, SF_HB = 8u; // SF_HB: a boundary strip of an overlapped halo exchange (carried to the flush of its parked column in bit 30 of py0)
	constexpr int PY_HB = 1<<30;
	for(;;) { // ---- tiles: strips as published by the producer, inside a strip x ascending
		mbar_wait_y(bar0+8u*s, ph, (lc.flags&LC_YIELD)!=0u);
		const bool first = xt==0u, last = xt+1u==tiles_x;
		if(first) { // ---- a new strip
			const uint32_t raw = tile_strip[s];
			if(raw==END) break;
			const uint32_t strip = raw&~STRIP_BND;
			y0 = (int)(strip%tiles_y)*TY; z0 = (int)(strip/tiles_y)*TZ;
			y = (uint32_t)y0+ly; z = (uint32_t)z0+lz;
			sf = ((y0==0||y0+TY>=(int)c.Ny||z0==0||z0+TZ>=(int)c.Nz) ? SF_BND : 0u)|((raw&STRIP_BND) ? SF_HB : 0u);
			if(y<c.Ny&&z<c.Nz&&!((c.Dy>1u&&(y==0u||y>=c.Ny-1u))||(c.Dz>1u&&(z==0u||z>=c.Nz-1u)))) sf |= SF_IN;
			if(has_zones&&zone_row_hit(c, y, z)) sf |= SF_ZONE;
			park_off = (uint32_t)CFG::BOX_BYTES+((kstrip&1u)*(uint32_t)CFG::ROWS+row)*(uint32_t)CFG::ES; // in stage 0, + box_off(b): this row's parked element of shifted box b
		}
		const bool bnd_yz = (sf&SF_BND)!=0u, in_yz = (sf&SF_IN)!=0u, zone_yz = (sf&SF_ZONE)!=0u;
		const bool wrap = s+1u==(uint32_t)S;
		const uint32_t s1 = wrap ? 0u : s+1u, ph1 = wrap ? ph^1u : ph, st1 = wrap ? sm0 : st+(uint32_t)CFG::STAGE_BYTES;
		if(!last&&(next_warp||bnd_yz)) mbar_wait_a(bar0+8u*s1, ph1);
		if(bnd_yz||(park&&xt<2u)) { // ---- rare: y/z wrap patches, the parked periodic-x column
			const int x0 = (int)xt*TX;
			if(bnd_yz) {
				if(first) patch_yz<CFG, true>(c, stage0+(size_t)s*CFG::STAGE_BYTES, x0, y0, z0, odd, tid, false);
				if(!last) patch_yz<CFG, true>(c, stage0+(size_t)s1*CFG::STAGE_BYTES, x0+TX, y0, z0, odd, tid, false);
				consumer_bar((uint32_t)NC);
			}
			if(park&&lx==lc.rowend_last) { // the thread that holds the row's last pair in the strip's last tile owns the periodic-x column of its row
				if(kstrip>0u&&xt==1u) { // the previous strip's column goes to global memory (its first tile has been written back: bar_head)
					mbar_wait_a(bar0+16u*(uint32_t)S, (kstrip-1u)&1u);
					flush_wrap<CFG>(c, stage0, (kstrip-1u)&1u, row, py0&~PY_HB, pz0, odd);
					if(py0&PY_HB) { __threadfence(); atomicAdd(c.bdone, 1u); } // one count per row of a boundary strip's parked column
				}
				if(first) { // park column 0 of the x-shifted boxes (pre-collision values; nothing else touches them before the strip's last tile)
#pragma unroll
					for(int b=0; b<Q; b++) if(box_shifted(b)) SP::ste(sm0+park_off+(uint32_t)CFG::box_off(b), SP::lde(st+row*(uint32_t)(TX*CFG::ES)+(uint32_t)CFG::box_off(b)));
				}
			}
		}
		if(in_yz) {
			const uint32_t bb = st+tid*(uint32_t)sizeof(R); // the pair's word in box b is at bb + box_off(b)
			const uint32_t fl2 = lds_u16(st+(uint32_t)CFG::FLAG_OFF+2u*tid);
			const uint32_t x = xt*(uint32_t)TX+lx;
			// ONE warp-uniform decision per tile: all 64 cells plain fluid or TYPE_E, no halo column, no column beyond the lattice -> fast body
			const uint32_t orfl = __reduce_or_sync(0xFFFFFFFFu, fl2); // every flag bit some cell of the warp carries: one reduction answers the three warp-uniform questions of a tile
			const uint32_t xbit = 1u<<(xt<31u ? xt : 31u);
			const bool slow = (orfl&~(E_BITS|T_BITS))!=0u||(edge_bits&xbit)!=0u;
			const uint32_t e2 = EQ ? fl2&E_BITS : 0u; // TYPE_E lanes (fast body)
			// element right of the pair's word in an x-shifted box: the next word of the row, or column 0 of the same row in the next stage / the parked column
			uint32_t nxt = bb+(uint32_t)sizeof(R);
			if(last ? end_last : end_full) nxt = !last ? st1+row*(uint32_t)(TX*CFG::ES) : park ? sm0+park_off : wrap_x ? st+row*(uint32_t)(TX*CFG::ES) : bb;
			const bool zone_warp = zone_yz||(zonex_bits&xbit)!=0u;
			if(EQ&&!last&&(xt+2u>=tiles_x||(orfl&~T_BITS)!=0u)) lean_prefetch_e<CFG>(c, lds_u16(st1+(uint32_t)CFG::FLAG_OFF+2u*tid), (uint64_t)x+(uint64_t)y*c.Px+(uint64_t)z*((uint64_t)c.Px*c.Ny));
			if(!slow) { // ---------------- fast body: 64 cells that all execute
				PairIn in;
				in.zones = false;
				in.any_e = EQ&&(orfl&E_BITS)!=0u; // warp-uniform: the selects of the TYPE_E lanes stay out of the common path
				in.e0 = (e2&0x00FFu)!=0u; in.e1 = (e2&0xFF00u)!=0u; in.n = (uint64_t)x+(uint64_t)y*c.Px+(uint64_t)z*((uint64_t)c.Px*c.Ny); // pure functions of live values: rematerialised where fast_prepare needs them. The TYPE_E lanes' rho / u are the boundary fields, loaded there (prefetched into L2 a tile ago)
				if(zone_warp) { // relaxation-zone data first: its global loads are in flight while the moments are accumulated
					const uint64_t n_row = (uint64_t)y*c.Px+(uint64_t)z*((uint64_t)c.Px*c.Ny);
					const ZoneRow zr = zone_row(c, y, z);
					in.nudge_vertical = c.nudge_vertical;
					in.zr0 = zone_cell(c, zr, x, n_row, (fl2&E_BITS&0x00FFu)==0u); // not for TYPE_E cells (FX/kernel.cpp:1524); the fast body only sees plain fluid and TYPE_E cells (+ TYPE_T on either)
					in.zr1 = zone_cell(c, zr, x+1u, n_row, (fl2&E_BITS&0xFF00u)==0u);
					in.zones = in.zr0.nudge||in.zr0.sponge||in.zr1.nudge||in.zr1.sponge;
				}
				Moments M;
				const f2 g0 = SP::dec(SP::ldw(bb));
				struct Raw { R wa, wb; E e0, e1; };
				const auto fetch = [&](const int k) -> Raw { // the pair's word of box A, and of box B -- or, in an x-shifted box B, the second element of its word and the element to the right of it
					Raw r;                                     // (element loads: the first element of the word belongs to the pair on the left, which may be writing it)
					r.wa = SP::ldw(bb+(uint32_t)CFG::box_off(1+2*k));
					if(pair_shifted(k)) { r.e0 = SP::lde(bb+(uint32_t)(CFG::box_off(2+2*k)+CFG::ES)); r.e1 = SP::lde(nxt+(uint32_t)CFG::box_off(2+2*k)); r.wb = r.wa; }
					else { r.wb = SP::ldw(bb+(uint32_t)CFG::box_off(2+2*k)); r.e0 = r.e1 = SP::low(r.wb); }
					return r;
				};
				const auto decode = [&](const int k, const Raw& r, f2& gi, f2& gj) { gi = SP::dec(r.wa); gj = SP::dec(pair_shifted(k) ? SP::join(r.e0, r.e1) : r.wb); };
				moments_of<SG>(g0, [&](const int k, f2& gi, f2& gj) { decode(k, fetch(k), gi, gj); }, M);
				FastK K;
				PairOut out;
				fast_prepare<FEAT>(c, a, in, M, scale, inv, K, out);
				if(TH) store_upre(c, (uint64_t)x+(uint64_t)y*c.Px+(uint64_t)z*((uint64_t)c.Px*c.Ny), (e2&0x00FFu)==0u, (e2&0xFF00u)==0u, out);
				if(UF) { // rho / u of the non-TYPE_E cells (UPDATE_FIELDS, FX/kernel.cpp:1709-1715)
					const uint64_t n = (uint64_t)x+(uint64_t)y*c.Px+(uint64_t)z*((uint64_t)c.Px*c.Ny);
					if(e2==0u) {
						*(float2*)(c.rho+n) = out.rho.v; *(float2*)(c.u+n) = out.ux.v; *(float2*)(c.u+c.N+n) = out.uy.v; *(float2*)(c.u+2ull*c.N+n) = out.uz.v;
					} else {
						if((e2&0x00FFu)==0u) { c.rho[n] = out.rho.v.x; c.u[n] = out.ux.v.x; c.u[c.N+n] = out.uy.v.x; c.u[2ull*c.N+n] = out.uz.v.x; }
						if((e2&0xFF00u)==0u) { c.rho[n+1ull] = out.rho.v.y; c.u[n+1ull] = out.ux.v.y; c.u[c.N+n+1ull] = out.uy.v.y; c.u[2ull*c.N+n+1ull] = out.uz.v.y; }
					}
				}
				// pass 2 re-reads the boxes (volatile shared loads: the DDFs do not stay in registers); the next group's words are fetched before this group's are stored
				const auto store = [&](const int k, const f2 gi, const f2 gj) { // f_i+1' goes to slot A, f_i' to slot B; whole words, no masks
					SP::stw(bb+(uint32_t)CFG::box_off(1+2*k), SP::enc(gj));
					if(pair_shifted(k)) SP::shift_out_both(bb+(uint32_t)CFG::box_off(2+2*k), nxt+(uint32_t)CFG::box_off(2+2*k), SP::enc(gi));
					else SP::stw(bb+(uint32_t)CFG::box_off(2+2*k), SP::enc(gi));
				};
				const auto axis = [&](const int ax, const Raw& r) { f2 gi, gj; decode(ax, r, gi, gj); fast_relax_axis(K, ax, gi, gj); store(ax, gi, gj); };
				const auto diag = [&](const int pl, const Raw& rp, const Raw& rm) {
					f2 gip, gjp, gim, gjm;
					decode(3+pl, rp, gip, gjp); decode(6+pl, rm, gim, gjm);
					fast_relax_diag(K, pl, gip, gjp, gim, gjm);
					store(3+pl, gip, gjp); store(6+pl, gim, gjm);
				};
				SP::stw(bb, SP::enc(fma2(K.omw, g0, K.g0add)));
				const Raw r0 = fetch(0), r1 = fetch(1);
				axis(0, r0);
				const Raw r2 = fetch(2);
				axis(1, r1);
				const Raw r3 = fetch(3), r6 = fetch(6);
				axis(2, r2);
				const Raw r4 = fetch(4), r7 = fetch(7);
				diag(0, r3, r6);
				const Raw r5 = fetch(5), r8 = fetch(8);
				diag(1, r4, r7);
				diag(2, r5, r8);
			} else lean_general_pair<CFG, FEAT>(c, a, bb, nxt, fl2, x, y, z, zone_warp, last&&(lc.flags&LC_ODD_X)!=0u&&lx==lc.rowend_last);
		}
		if(bnd_yz) { // write the wrapped elements back to their real addresses (TMA clips them)
			consumer_bar((uint32_t)NC);
			patch_yz<CFG, false>(c, stage0+(size_t)s*CFG::STAGE_BYTES, (int)xt*TX, y0, z0, odd, tid, park);
		}
		fence_async_smem(); // make this thread's shared-memory writes visible to the TMA store
		__syncwarp();
		if((tid&31u)==0u) mbar_arrive_a(bar0+8u*((uint32_t)S+s)); // one arrival per warp
		s = s1; ph = ph1; st = st1;
		if(last) { xt = 0u; py0 = y0|((sf&SF_HB) ? PY_HB : 0); pz0 = z0; kstrip++; } else xt++;
	}
	if(park&&kstrip>0u&&lx==lc.rowend_last) { // the last strip's periodic-x column
		mbar_wait_a(bar0+16u*(uint32_t)S, (kstrip-1u)&1u);
		flush_wrap<CFG>(c, stage0, (kstrip-1u)&1u, row, py0&~PY_HB, pz0, odd);
		if(py0&PY_HB) { __threadfence(); atomicAdd(c.bdone, 1u); }
	}
}

} // anonymous namespace
} // namespace luw
