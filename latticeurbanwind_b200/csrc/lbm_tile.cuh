// stream_collide as a persistent, TMA-staged, warp-specialised sm_100a kernel (re-implements FX/kernel.cpp:1475-1780).
//
// Data movement. The Esoteric-Pull step (FX/kernel.cpp:1338-1351) makes every cell read and then overwrite the SAME 19 DDF slots:
// 10 in its own cell and 9 in the neighbours n+c_i (i odd). For a box of cells, the slots of direction pair (i,i+1) therefore form
// two boxes of the SoA array: slot A = (t odd ? i : i+1) at the tile origin, and slot B = (t odd ? i+1 : i) at origin + c_i. The
// kernel moves these 19 boxes (+ the flag box) with cp.async.bulk.tensor (TMA): global -> shared, collide in place in shared memory,
// shared -> global. Boxes are shifted by whole rows / planes for the y and z components of c_i, so every thread finds its values at
// the same offset of 19 dense shared-memory boxes (conflict-free 32/64-bit LDS with immediate offsets) and the LSU never issues a
// global DDF access. The innermost TMA coordinate must be 16-byte aligned (measured: an odd x origin raises "illegal instruction" on
// B200), so the five x+1 shifted streams are loaded at the tile's own x origin and read one element to the right; the element that
// belongs to the tile's last column then sits in column 0 of the NEXT tile's box. A CTA therefore walks whole x-strips, tile after
// tile, and finishes the last column of a tile inside the already prefetched next stage.
// Periodic x (Dx == 1): the +x neighbour of the lattice's last column is column 0, i.e. column 0 of the strip's FIRST tile, which
// has long been written back when the strip's last tile is processed. Its five elements per row are therefore parked in shared memory
// when the first tile passes (pads behind the five shifted boxes of stage 0), read and updated there by the last tile, and written
// to global memory with plain stores one strip later, after the producer has seen the first tile's TMA store complete.
// Cells that do not execute (solid / gas / halo, FX/kernel.cpp:1486-1490) leave their slots untouched in shared memory, so the
// write-back of the box is the identity for them -- legal because no other cell touches those slots.
//
// Pipeline. One CTA = TILE/2 consumer threads (two x-adjacent cells per thread, packed FP32x2 arithmetic, lbm_vec.cuh) + one producer
// warp, over a ring of STAGES shared-memory stages with mbarriers:  producer: expect_tx + 20 TMA loads -> full[s];  consumers: wait
// full[s] -> collide in place -> fence.proxy.async -> arrive done[s];  producer: wait done[s] -> 19 TMA stores -> commit ->
// wait_group.read -> refill the stage with the tile STAGES ahead. CTAs are persistent (grid = resident CTAs); strips are handed out
// dynamically through an atomic counter (SMs do not all see the same memory bandwidth -- a static split leaves the fast ones idle at the end);
// the producer publishes the strip id of every stage in shared memory, an END id tells the consumers to leave.
//
// TYPE_E cells (FX/kernel.cpp:1503-1515,1747) take their rho/u from the boundary fields and set f := feq; they do not depend on the
// streamed DDFs at all. They are kept out of the packed main path: lanes that hold a TYPE_E cell are overwritten afterwards by a
// scalar equilibrium, and their rho/u are prefetched into L2 one tile ahead.
//
// Periodic wrap in y / z. TMA zero-fills the part of a box that lies outside the lattice and clips it on the way back. In strips that
// touch the y / z boundary the consumers patch those elements from / to their wrapped addresses (FX/kernel.cpp:920-931) with plain loads and stores.
#pragma once
#include <cuda.h>
#include "lbm_launch.h"
#include "lbm_vec.cuh"

namespace luw {
namespace {

// ------------------------------------------------------------------ PTX wrappers: mbarrier, TMA, proxy fence
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, const uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, const uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, const uint32_t parity) {
	uint32_t done;
	asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(b)), "r"(parity), "r"(1000000u) : "memory"); // suspend-time hint [ns]
	return done!=0u;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, const uint32_t parity) {
	if(mbar_try(b, parity)) return;
	uint32_t spins = 0u;
	while(!mbar_try(b, parity)) if(++spins>(1u<<24)) __trap(); // a lost arrival must abort the launch, not hang the device
}
// producer-side wait: polls with a growing sleep in between, so that the idle warp does not take issue slots from the consumers
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* b, const uint32_t parity) {
	uint32_t ns = 64u, spins = 0u;
	while(!mbar_try(b, parity)) {
		__nanosleep(ns);
		if(ns<1024u) ns <<= 1;
		if(++spins>(1u<<22)) __trap();
	}
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, const int c0, const int c1, const int c2, const int c3) {
	asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
		:: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, const int c0, const int c1, const int c2) {
	asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
		:: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, const int c0, const int c1, const int c2, const int c3) {
	asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
		:: "l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); } // all but the most recent group have been read from shared memory
__device__ __forceinline__ void tma_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// all but the `n` most recent groups complete (the operand is an immediate; a smaller n than asked for only waits for more)
__device__ __forceinline__ void tma_wait_all_but(const uint32_t n) {
	switch(n<7u ? n : 7u) {
		case 0u: asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); break;
		case 1u: asm volatile("cp.async.bulk.wait_group 1;" ::: "memory"); break;
		case 2u: asm volatile("cp.async.bulk.wait_group 2;" ::: "memory"); break;
		case 3u: asm volatile("cp.async.bulk.wait_group 3;" ::: "memory"); break;
		case 4u: asm volatile("cp.async.bulk.wait_group 4;" ::: "memory"); break;
		case 5u: asm volatile("cp.async.bulk.wait_group 5;" ::: "memory"); break;
		case 6u: asm volatile("cp.async.bulk.wait_group 6;" ::: "memory"); break;
		default: asm volatile("cp.async.bulk.wait_group 7;" ::: "memory"); break;
	}
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void consumer_bar(const uint32_t nthreads) { asm volatile("bar.sync 1, %0;" :: "r"(nthreads) : "memory"); }
// LUW_PRODUCER_NAMED_BARRIER (experiment, off): named barriers 2 .. 2+STAGES-1 carry "stage s has been collided": the consumers arrive (non-blocking), the
// producer warp syncs (blocks in hardware without taking issue slots; polling an mbarrier costs the producer ~12 % of all issued instructions,
// profiles/r1_ncu_urban_fp16s.md). Measured slower than polling on the channel case (61.9 vs 68.0 GLUP/s with the single-pass kernel): BAR.ARV drains the
// consumers' pending shared-memory stores, and the dynamic barrier id makes ptxas reserve all 16 barriers.
__device__ __forceinline__ void bar_arrive_id(const uint32_t id, const uint32_t nthreads) { asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void bar_sync_id(const uint32_t id, const uint32_t nthreads) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ bool elect_one() { // one lane of the (converged) warp
	uint32_t p;
	asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(p));
	return p!=0u;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

#ifdef LUW_TRACE // development aid: per-tile timestamps of CTA `LUW_TRACE` (producer: done seen / stores committed / smem read / loads issued; consumer warp 0: start / end)
__device__ long long g_trace[6][2048];
#define TRACE(slot, q) do { if(blockIdx.x==(LUW_TRACE)&&(q)<2040u) g_trace[slot][q] = clock64(); } while(0)
__device__ __forceinline__ long long gtimer() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define TRACE_CLK(i) do { if(blockIdx.x==(LUW_TRACE)&&threadIdx.x==0u) { g_trace[i][2047] = gtimer(); g_trace[i+1][2047] = clock64(); } } while(0)
#else
#define TRACE_CLK(i) do {} while(0)
#define TRACE(slot, q) do {} while(0)
#endif

// ------------------------------------------------------------------ geometry of a tile
// Box b of a stage: b = 0 is f0; b = 1+2k is slot A of pair k (own cell: holds f_i, receives f_i+1); b = 2+2k is slot B of pair k
// (neighbour n+c_i: holds f_i+1, receives f_i). In shared memory the nine A boxes are contiguous (they are moved by ONE TMA operation that
// walks the slot dimension with element stride 2), followed by the nine B boxes. Pairs 0,3,4,6,7 have c_x = +1: their B boxes are the
// x-shifted ones; each is followed by a 128-byte pad (stage 0's pads park the periodic-x column, see above).
__host__ __device__ constexpr bool pair_shifted(const int k) { return k==0||k==3||k==4||k==6||k==7; }
__host__ __device__ constexpr bool box_shifted(const int b) { return b>0&&(b&1)==0&&pair_shifted((b-2)/2); }
__host__ __device__ constexpr int pads_before_pair(const int k) { return (k>0)+(k>3)+(k>4)+(k>6)+(k>7); }
template<int P_, int TX_, int TY_, int TZ_, int STAGES_, int CTAS_, bool TWOPASS_ = false, bool LEAN_ = false> struct TileCfg {
	static constexpr int P = P_, TX = TX_, TY = TY_, TZ = TZ_, STAGES = STAGES_, CTAS_PER_SM = CTAS_;
	static constexpr bool TWOPASS = TWOPASS_; // FAST arithmetic only: collide in two passes over the shared-memory boxes (fewer registers, more resident CTAs)
	static constexpr bool LEAN = LEAN_; // FAST arithmetic only: k_stream_collide_lean (warp-uniform fast path for plain fluid cells) instead of k_stream_collide_tile
	static constexpr int TILE = TX*TY*TZ, ROWS = TY*TZ, CONSUMERS = TILE/2, THREADS = CONSUMERS+32;
	static constexpr int ES = (P_==P_FP32) ? 4 : 2;
	static constexpr int BOX_BYTES = TILE*ES, PAD = 128;
	static constexpr int FLAG_OFF = Q*BOX_BYTES+5*PAD;
	static constexpr int STAGE_BYTES = FLAG_OFF+TILE; // 19 DDF boxes + 5 pads + the flag box
	static constexpr int LOAD_BYTES = Q*BOX_BYTES+TILE; // what one stage's TMA loads deliver
	static constexpr int SMEM_BYTES = STAGES*STAGE_BYTES+(2*STAGES+1)*8+STAGES*12+4+128; // + mbarriers, + the strip id of every stage, + slack for 128 B alignment of the first stage
	__host__ __device__ static constexpr int box_off(const int b) {
		return b==0 ? 0 : (b&1) ? (1+(b-1)/2)*BOX_BYTES : (10+(b-2)/2)*BOX_BYTES+PAD*pads_before_pair((b-2)/2);
	}
	static_assert(TX%64==0&&TX<=256, "rows are whole warps; TMA boxes are at most 256 elements wide");
	static_assert(CONSUMERS%32==0&&BOX_BYTES%128==0&&TILE%128==0&&2*ROWS*ES<=PAD, "tile shape");
	static_assert(STAGES<=14, "one named barrier per stage (ids 2..15)");
};
// register budget per thread for `ctas` resident CTAs of `warps` warps: the register file is partitioned per SM sub-partition (4 x 16384), a CTA's warps are
// dealt round-robin to the four, so ceil(warps*ctas/4) warps must fit into 16384 registers, in units of 8 per thread (measured: 5 CTAs x 5 warps at 80
// registers gave 4 resident CTAs, 3 x 5 at 136 gave 2)
__host__ __device__ constexpr int tile_max_regs(const int warps, const int ctas) { const int r = (16384/((warps*ctas+3)/4))/32/8*8; return r>128 ? 128 : r; }
// pair k = (i-1)/2, i odd: c_i
__device__ __forceinline__ void pair_shift(const int k, int& cx, int& cy, int& cz) {
	const int CXv[9] = {1,0,0,1,1,0, 1, 1, 0}, CYv[9] = {0,1,0,1,0,1,-1, 0, 1}, CZv[9] = {0,0,1,0,1,1, 0,-1,-1};
	cx = CXv[k]; cy = CYv[k]; cz = CZv[k];
}

template<int P> struct PairCodec;
template<> struct PairCodec<P_FP32> { // two floats
	typedef float2 R;
	typedef float E;
	static __device__ __forceinline__ f2 dec(const R r) { f2 v; v.v = r; return v; }
	static __device__ __forceinline__ R enc(const f2 v) { return v.v; }
	static __device__ __forceinline__ R mix(const bool k0, const bool k1, const R n, const R o) { return make_float2(k0 ? n.x : o.x, k1 ? n.y : o.y); }
	static __device__ __forceinline__ E low(const R w) { return w.x; }
	// x+1 shifted stream: elements (lx+1, lx+2) from the thread's own word w0 = (lx, lx+1) and the element right of it
	static __device__ __forceinline__ R shift_in(const R w0, const uint8_t* next) { return make_float2(w0.y, *(const float*)next); }
	static __device__ __forceinline__ void shift_out(uint8_t* own, uint8_t* next, const R n, const bool k0, const bool k1) { if(k0) ((float*)own)[1] = n.x; if(k1) *(float*)next = n.y; }
	static __device__ __forceinline__ void shift_out_both(uint8_t* own, uint8_t* next, const R n) { ((float*)own)[1] = n.x; *(float*)next = n.y; }
	// odd Nx, the row's last pair: cell 0 IS the last column (its +x element is `next`), cell 1 does not exist
	static __device__ __forceinline__ R shift_in_odd(const uint8_t* next) { return make_float2(*(const float*)next, 0.0f); }
	// ... by element loads: the FIRST element of the pair's word belongs to the pair on the left, which may be writing it (a whole-word load is a data race, if a harmless one)
	static __device__ __forceinline__ R shift_in_e(const uint8_t* own, const uint8_t* next) { return make_float2(((const float*)own)[1], *(const float*)next); }
	static __device__ __forceinline__ void shift_out_odd(uint8_t* next, const R n) { *(float*)next = n.x; }
	typedef uint32_t M; // lane mask of a pair: bit 0 / bit 1 = cell 0 / 1 takes the new value
	static __device__ __forceinline__ M mask(const bool k0, const bool k1) { return (k0 ? 1u : 0u)|(k1 ? 2u : 0u); }
	static __device__ __forceinline__ R mixm(const M m, const R n, const R o) { return make_float2((m&1u) ? n.x : o.x, (m&2u) ? n.y : o.y); }
};
template<> struct PairCodec<P_FP16S> { // half2 holding 2^15 f; the 2^15 is folded into the FAST collision, applied in STRICT
	typedef uint32_t R;
	typedef uint16_t E;
	static __device__ __forceinline__ f2 dec_raw(const R r) { f2 v; v.v = __half22float2(*reinterpret_cast<const __half2*>(&r)); return v; }
	static __device__ __forceinline__ R enc_raw(const f2 v) { const __half2 h = __float22half2_rn(v.v); return *reinterpret_cast<const R*>(&h); }
	static __device__ __forceinline__ f2 dec(const R r) { return sm(3.0517578E-5f, dec_raw(r)); } // exact scaling; scalar so that it is never fused into a later add
	static __device__ __forceinline__ R enc(const f2 v) { return enc_raw(sm(32768.0f, v)); }
	static __device__ __forceinline__ R mix(const bool k0, const bool k1, const R n, const R o) { return (k0 ? (n&0xFFFFu) : (o&0xFFFFu))|(k1 ? (n&0xFFFF0000u) : (o&0xFFFF0000u)); }
	static __device__ __forceinline__ E low(const R w) { return (uint16_t)(w&0xFFFFu); }
	static __device__ __forceinline__ R shift_in(const R w0, const uint8_t* next) { return __byte_perm(w0, (uint32_t)*(const uint16_t*)next, 0x5432); }
	static __device__ __forceinline__ void shift_out(uint8_t* own, uint8_t* next, const R n, const bool k0, const bool k1) { if(k0) ((uint16_t*)own)[1] = (uint16_t)(n&0xFFFFu); if(k1) *(uint16_t*)next = (uint16_t)(n>>16); }
	static __device__ __forceinline__ void shift_out_both(uint8_t* own, uint8_t* next, const R n) { ((uint16_t*)own)[1] = (uint16_t)(n&0xFFFFu); *(uint16_t*)next = (uint16_t)(n>>16); }
	static __device__ __forceinline__ R shift_in_odd(const uint8_t* next) { return (uint32_t)*(const uint16_t*)next; }
	static __device__ __forceinline__ R shift_in_e(const uint8_t* own, const uint8_t* next) { return (uint32_t)((const uint16_t*)own)[1]|((uint32_t)*(const uint16_t*)next<<16); }
	static __device__ __forceinline__ void shift_out_odd(uint8_t* next, const R n) { *(uint16_t*)next = (uint16_t)(n&0xFFFFu); }
	typedef uint32_t M; // bit mask of a pair: the halves that take the new value
	static __device__ __forceinline__ M mask(const bool k0, const bool k1) { return (k0 ? 0x0000FFFFu : 0u)|(k1 ? 0xFFFF0000u : 0u); }
	static __device__ __forceinline__ R mixm(const M m, const R n, const R o) { return (n&m)|(o&~m); }
};
template<> struct PairCodec<P_FP16C> {
	typedef uint32_t R;
	typedef uint16_t E;
	static __device__ __forceinline__ f2 dec(const R r) { // both halves at once, see Ddf<P_FP16C>::dec
		f2 m = mk2(__uint_as_float((r<<12)&0x07FFF000u), __uint_as_float((r>>4)&0x07FFF000u));
		m.v = __fmul2_rn(m.v, make_float2(5.192296858534828e33f, 5.192296858534828e33f)); // 2^112, exact
		return mk2(__uint_as_float(__float_as_uint(m.v.x)|((r<<16)&0x80000000u)), __uint_as_float(__float_as_uint(m.v.y)|(r&0x80000000u)));
	}
	static __device__ __forceinline__ R enc(const f2 v) { return (uint32_t)Ddf<P_FP16C>::enc(v.v.x)|((uint32_t)Ddf<P_FP16C>::enc(v.v.y)<<16); }
	// FAST encode: f * 2^-112 has the FP16C exponent in a float's exponent field, for subnormal codes too (the product is then a subnormal float); adding
	// half an FP16C ulp (0x800) and dropping 12 bits is the reference's rounding (FX/kernel.cpp:870-875), including its wrap-around above the format's range.
	// Differs from the exact encoder only through the multiplication's own rounding of subnormal results, 12 bits below the FP16C ulp (a double rounding
	// that moves a result by one code when it lands exactly on a tie).
	static __device__ __forceinline__ R enc_fast(const f2 v) {
		const float2 g = __fmul2_rn(v.v, make_float2(1.925929944387236e-34f, 1.925929944387236e-34f)); // 2^-112
		const uint32_t bx = __float_as_uint(g.x), by = __float_as_uint(g.y);
		const uint32_t lo = (((bx+0x800u)>>12)&0x7FFFu)|((bx>>16)&0x8000u);
		const uint32_t hi = (((by+0x800u)<<4)&0x7FFF0000u)|(by&0x80000000u);
		return lo|hi;
	}
	static __device__ __forceinline__ R mix(const bool k0, const bool k1, const R n, const R o) { return (k0 ? (n&0xFFFFu) : (o&0xFFFFu))|(k1 ? (n&0xFFFF0000u) : (o&0xFFFF0000u)); }
	static __device__ __forceinline__ E low(const R w) { return (uint16_t)(w&0xFFFFu); }
	static __device__ __forceinline__ R shift_in(const R w0, const uint8_t* next) { return __byte_perm(w0, (uint32_t)*(const uint16_t*)next, 0x5432); }
	static __device__ __forceinline__ void shift_out(uint8_t* own, uint8_t* next, const R n, const bool k0, const bool k1) { if(k0) ((uint16_t*)own)[1] = (uint16_t)(n&0xFFFFu); if(k1) *(uint16_t*)next = (uint16_t)(n>>16); }
	static __device__ __forceinline__ void shift_out_both(uint8_t* own, uint8_t* next, const R n) { ((uint16_t*)own)[1] = (uint16_t)(n&0xFFFFu); *(uint16_t*)next = (uint16_t)(n>>16); }
	static __device__ __forceinline__ R shift_in_odd(const uint8_t* next) { return (uint32_t)*(const uint16_t*)next; }
	static __device__ __forceinline__ R shift_in_e(const uint8_t* own, const uint8_t* next) { return (uint32_t)((const uint16_t*)own)[1]|((uint32_t)*(const uint16_t*)next<<16); }
	static __device__ __forceinline__ void shift_out_odd(uint8_t* next, const R n) { *(uint16_t*)next = (uint16_t)(n&0xFFFFu); }
	typedef uint32_t M; // bit mask of a pair: the halves that take the new value
	static __device__ __forceinline__ M mask(const bool k0, const bool k1) { return (k0 ? 0x0000FFFFu : 0u)|(k1 ? 0xFFFF0000u : 0u); }
	static __device__ __forceinline__ R mixm(const M m, const R n, const R o) { return (n&m)|(o&~m); }
};

// ------------------------------------------------------------------ boundary helpers
// Boxes the TMA store must not write: (a) negative origin (measured: illegal instruction for stores; loads zero-fill), (b) a row/plane
// shifted by -1 in a partial tile, where the last lattice row/plane would be rewritten on behalf of cells that do not exist while its
// real owners (row/plane 0, through the periodic wrap) update it from another strip. The consumers write such boxes back (patch_yz).
template<class CFG> __device__ __forceinline__ bool box_by_threads(const DomainConst& c, const int cy, const int cz, const int y0, const int z0) {
	return y0+cy<0||z0+cz<0||(cy<0&&y0+CFG::TY>(int)c.Ny)||(cz<0&&z0+CFG::TZ>(int)c.Nz);
}
// y/z-wrap patch of stage `st` (tile origin x0,y0,z0): IN = fill the zero-filled elements from their wrapped addresses, !IN = write them back.
// `skip_x0`: column x = 0 of the x-shifted boxes is written by the periodic-x flush instead (see the head comment).
template<class CFG, bool IN> __device__ __noinline__ void patch_yz(const DomainConst& c, uint8_t* st, const int x0, const int y0, const int z0, const uint32_t odd, const uint32_t tid, const bool skip_x0) {
	typedef typename Ddf<CFG::P>::T T;
	constexpr int TX = CFG::TX, TY = CFG::TY;
	T* const fi = (T*)c.fi;
	const uint64_t rowN = c.Px, planeN = (uint64_t)c.Px*c.Ny;
	for(int k=1; k<9; k++) { // pair 0 (+x) has no y/z shift
		int cx, cy, cz; pair_shift(k, cx, cy, cz);
		const uint32_t slot = odd ? 2u*k+2u : 2u*k+1u;
		T* box = (T*)(st+CFG::box_off(2+2*k));
		const bool whole = box_by_threads<CFG>(c, cy, cz, y0, z0); // loaded by TMA, but written back here in full
		for(uint32_t e=tid; e<(uint32_t)CFG::TILE; e+=(uint32_t)CFG::CONSUMERS) {
			const int x = x0+(int)(e%TX), y = y0+(int)((e/TX)%TY), z = z0+(int)(e/(TX*TY));
			if(x>=(int)c.Nx||y>=(int)c.Ny||z>=(int)c.Nz) continue; // no cell owns this element
			const int yn = y+cy, zn = z+cz;
			if(yn>=0&&yn<(int)c.Ny&&zn>=0&&zn<(int)c.Nz&&(IN||!whole)) continue; // inside the lattice: TMA moves it
			if(!IN&&skip_x0&&cx!=0&&x==0) continue;
			const uint64_t nn = (uint64_t)x+(uint64_t)(yn<0 ? (int)c.Ny-1 : yn>=(int)c.Ny ? 0 : yn)*rowN+(uint64_t)(zn<0 ? (int)c.Nz-1 : zn>=(int)c.Nz ? 0 : zn)*planeN;
			if(IN) box[e] = fi[(uint64_t)slot*c.N+nn]; else fi[(uint64_t)slot*c.N+nn] = box[e];
		}
	}
}
// periodic-x flush: the parked column-0 elements of the strip with tile origin (y0,z0) go to global memory (row `row` of the tile)
template<class CFG> __device__ __noinline__ void flush_wrap(const DomainConst& c, const uint8_t* stage0, const uint32_t par, const uint32_t row, const int y0, const int z0, const uint32_t odd) {
	typedef typename Ddf<CFG::P>::T T;
	const int y = y0+(int)(row%(uint32_t)CFG::TY), z = z0+(int)(row/(uint32_t)CFG::TY);
	if(y>=(int)c.Ny||z>=(int)c.Nz) return;
	T* const fi = (T*)c.fi;
	const uint64_t rowN = c.Px, planeN = (uint64_t)c.Px*c.Ny;
	for(int k=0; k<9; k++) {
		int cx, cy, cz; pair_shift(k, cx, cy, cz);
		if(cx==0) continue;
		const uint32_t slot = odd ? 2u*k+2u : 2u*k+1u;
		const int yn = y+cy, zn = z+cz;
		const uint64_t nn = (uint64_t)(yn<0 ? (int)c.Ny-1 : yn>=(int)c.Ny ? 0 : yn)*rowN+(uint64_t)(zn<0 ? (int)c.Nz-1 : zn>=(int)c.Nz ? 0 : zn)*planeN; // x = 0
		fi[(uint64_t)slot*c.N+nn] = *(const T*)(stage0+CFG::box_off(2+2*k)+CFG::BOX_BYTES+(par*(uint32_t)CFG::ROWS+row)*(uint32_t)CFG::ES);
	}
}

// TYPE_E cell: rho/u are boundary data; Coriolis shift, clamp, f := feq (FX/kernel.cpp:1503-1522,1686-1716,1747). Scalar, in the reference's
// operation order (relaxation zones never apply to TYPE_E cells, FX/kernel.cpp:1524). feq[] in units of `scale`.
template<uint32_t FEAT> __device__ __forceinline__ void equilibrium_cell(const DomainConst& c, const StepArgs& a, const float rhon, float uxn, float uyn, float uzn, const float scale, float* feq) {
	constexpr bool VF = (FEAT&F_VOLUME_FORCE)!=0u;
	if(VF) {
		float fxn, fyn, fzn;
		luw_force(c, a, 0u, 0u, 0u, TYPE_E, false, rhon, uxn, uyn, uzn, fxn, fyn, fzn);
		const float rho2 = 0.5f/rhon;
		uxn = clampc(fmaf(fxn, rho2, uxn)); uyn = clampc(fmaf(fyn, rho2, uyn)); uzn = clampc(fmaf(fzn, rho2, uzn));
	} else {
		uxn = clampc(uxn); uyn = clampc(uyn); uzn = clampc(uzn);
	}
	f_eq(rhon, uxn, uyn, uzn, feq);
	if(scale!=1.0f) {
#pragma unroll
		for(int i=0; i<Q; i++) feq[i] *= scale; // exact (power of two)
	}
}

// thermal step in two kernels (DomainConst::upre): the pair's velocity before the force half-step; w0 / w1: the cell executes and is not TYPE_E (k_thermal_g takes u of TYPE_E cells from the boundary field)
__device__ __forceinline__ void store_upre(const DomainConst& c, const uint64_t n, const bool w0, const bool w1, const PairOut& o) {
	if(w0&&w1) { *(float2*)(c.upre+n) = o.upx.v; *(float2*)(c.upre+c.N+n) = o.upy.v; *(float2*)(c.upre+2ull*c.N+n) = o.upz.v; }
	else {
		if(w0) { c.upre[n] = o.upx.v.x; c.upre[c.N+n] = o.upy.v.x; c.upre[2ull*c.N+n] = o.upz.v.x; }
		if(w1) { c.upre[n+1ull] = o.upx.v.y; c.upre[c.N+n+1ull] = o.upy.v.y; c.upre[2ull*c.N+n+1ull] = o.upz.v.y; }
	}
}
// rho / u of a strip's west-face cells (x = 0; TYPE_E in every open-boundary case) into L2: by the producer warp when it issues the loads of the strip's first tile (it is
// the one that knows the strip; the consumers used to peek at the next stage's strip id for this, which the race checker rightly flags). One sector per lane.
template<class CFG> __device__ __forceinline__ void prefetch_west_face(const DomainConst& c, const uint32_t lane, const int y0, const int z0) {
	const uint32_t r = lane>>2, a = lane&3u;
	if(r<(uint32_t)CFG::ROWS) {
		const uint32_t y = (uint32_t)y0+r%(uint32_t)CFG::TY, z = (uint32_t)z0+r/(uint32_t)CFG::TY;
		if(y<c.Ny&&z<c.Nz) { const uint64_t m = (uint64_t)y*c.Px+(uint64_t)z*((uint64_t)c.Px*c.Ny); prefetch_l2(a==0u ? c.rho+m : c.u+(uint64_t)(a-1u)*c.N+m); }
	}
}

// ------------------------------------------------------------------ the kernel
template<class CFG, uint32_t FEAT, bool FAST> __global__ void __maxnreg__(((FAST&&CFG::TWOPASS) ? tile_max_regs(CFG::THREADS/32, CFG::CTAS_PER_SM) : 128)) // the single-pass paths hold all 19 DDF pairs: 128 registers, residency as it comes
k_stream_collide_tile(const __grid_constant__ DomainConst c, const __grid_constant__ StepArgs a, const __grid_constant__ TileMaps maps, const uint32_t tiles_x, const uint32_t tiles_y, const uint32_t tiles_z, const uint32_t opts) {
	constexpr int P = CFG::P, TX = CFG::TX, TY = CFG::TY, TZ = CFG::TZ, TILE = CFG::TILE, S = CFG::STAGES, NC = CFG::CONSUMERS;
	typedef PairCodec<P> PC;
	typedef typename PC::R R;
	typedef typename PC::E E;
	constexpr bool UF = (FEAT&F_UPDATE_FIELDS)!=0u, EQ = (FEAT&F_EQUILIBRIUM)!=0u, VF = (FEAT&F_VOLUME_FORCE)!=0u, TH = (FEAT&F_TEMPERATURE)!=0u;

	extern __shared__ uint8_t smem_raw[];
	uint8_t* const stage0 = smem_raw+((128u-(smem_u32(smem_raw)&127u))&127u); // 128 B aligned; derived from the __shared__ symbol so that LDS/STS are emitted
	uint64_t* const bar_full = (uint64_t*)(stage0+(size_t)S*CFG::STAGE_BYTES);
	uint64_t* const bar_done = bar_full+S;
	uint64_t* const bar_head = bar_done+S;
	volatile uint32_t* const tile_strip = (volatile uint32_t*)(bar_head+1); // strip id of the tile in each stage
	volatile int* const tile_yz = (volatile int*)(tile_strip+S); // its y0, z0 (producer only: saves the store side two integer divisions per tile)

	const uint32_t tid = threadIdx.x;
	TRACE_CLK(0);
	if(tid==0u) {
		if(blockIdx.x==0u) c.sched[(uint32_t)(a.t&1ull)^1u] = 0u; // strip counter of the NEXT step (steps alternate between two counters, so no memset node sits between launches)
		for(int s=0; s<S; s++) { mbar_init(bar_full+s, 1u); mbar_init(bar_done+s, (uint32_t)(NC/32)); } // done: one arrival per consumer WARP
		mbar_init(bar_head, 1u);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	// tile sequence of this CTA: strips (one (y,z) tile row each) drawn from the counter c.sched[t&1] (zeroed by the previous step's launch, or by the host when the parity repeats); inside a strip x ascending
	const uint32_t nstrips = tiles_y*tiles_z;
	constexpr uint32_t END = 0xFFFFFFFFu, STRIP_BND = 0x80000000u; // tile_strip[s]: strip id | STRIP_BND for a boundary strip of an overlapped halo exchange (DomainConst::so_nb)
	const uint32_t odd = (uint32_t)(a.t&1ull);
	const bool wrap_x = c.Dx==1u; // the lattice is periodic in x inside this domain
	const bool park = wrap_x&&tiles_x>=2u; // ... and the wrapped column lives in another tile than the last one

	if(tid>=(uint32_t)NC) { // ---------------------------------------------------------------- producer warp: TMA loads and stores
		// The whole warp walks the loops (uniform control flow keeps coordinates and addresses in uniform registers); lane 0 issues.
		const bool leader = (tid&31u)==0u;
		uint32_t lstrip = 0u, lxt = 0u, issued = 0u, lbnd = 0u; // tile the next load belongs to; tiles whose loads have been issued
		bool ended = false;
		const bool lag = (opts&1u)!=0u; // refill the stage of the PREVIOUS tile (its stores were committed a tile-time ago) instead of waiting for this tile's stores to be read
		const auto issue_loads = [&]() {
			const int s = (int)(issued%(uint32_t)S);
			if(lxt==0u) { // next strip
				uint32_t v = 0u;
				if(leader) v = atomicAdd(c.sched+odd, 1u); // the counter of this step parity; the other one is being zeroed for the next step
				lstrip = __shfl_sync(0xFFFFFFFFu, v, 0);
				if(lstrip>=nstrips) { // no more work: tell the consumers
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
				tile_strip[s] = lstrip|lbnd; // published by the barrier's release / acquire
				tile_yz[2*s] = y0; tile_yz[2*s+1] = z0;
				mbar_expect_tx(bar_full+s, (uint32_t)CFG::LOAD_BYTES);
				tma_load_3d(st+CFG::FLAG_OFF, &maps.flags, bar_full+s, x0, y0, z0);
				tma_load_4d(st, &maps.fi, bar_full+s, x0, y0, z0, 0);
				tma_load_4d(st+CFG::box_off(1), &maps.fiA, bar_full+s, x0, y0, z0, odd ? 1 : 2); // the nine A slots: odd ? 1,3,..,17 : 2,4,..,18
#pragma unroll
				for(int k=0; k<9; k++) {
					int cx, cy, cz; pair_shift(k, cx, cy, cz);
					const int i = 2*k+1;
					tma_load_4d(st+CFG::box_off(2+2*k), &maps.fi, bar_full+s, x0, y0+cy, z0+cz, odd ? i+1 : i); // x shift applied by the readers
				}
			}
			if(++lxt==tiles_x) lxt = 0u;
			issued++;
		};
		for(int i=0; i<S&&!ended; i++) issue_loads();
		uint32_t sxt = 0u; // x tile of the next store
		for(uint32_t q=0u; q<issued; q++) { // `issued` keeps growing until the strips run out
			const int s = (int)(q%(uint32_t)S);
#ifndef LUW_PRODUCER_NAMED_BARRIER
			mbar_wait_backoff(bar_done+s, (q/(uint32_t)S)&1u);
#else
			bar_sync_id(2u+(uint32_t)s, (uint32_t)CFG::THREADS); // all consumers have arrived: stage s is collided and fenced for the async proxy
#endif
			if(leader) TRACE(0, q);
			const int x0 = (int)sxt*TX, y0 = tile_yz[2*s], z0 = tile_yz[2*s+1];
			const uint8_t* st = stage0+(size_t)s*CFG::STAGE_BYTES;
			const bool inner = y0>0&&z0>0&&y0+TY<=(int)c.Ny&&z0+TZ<=(int)c.Nz; // every box lies inside the lattice: all 19 go through TMA
			const bool last_of_strip = sxt+1u==tiles_x;
			if(leader) {
				tma_store_4d(&maps.fi, st, x0, y0, z0, 0);
				tma_store_4d(&maps.fiA, st+CFG::box_off(1), x0, y0, z0, odd ? 1 : 2);
#pragma unroll
				for(int k=0; k<9; k++) {
					int cx, cy, cz; pair_shift(k, cx, cy, cz);
					const int i = 2*k+1;
					if(inner||!box_by_threads<CFG>(c, cy, cz, y0, z0)) tma_store_4d(&maps.fi, st+CFG::box_off(2+2*k), x0, y0+cy, z0+cz, odd ? i+1 : i);
				}
				tma_commit();
				if(last_of_strip&&(tile_strip[s]&STRIP_BND)!=0u) { tma_wait_all0(); __threadfence(); atomicAdd(c.bdone, 1u); } // a boundary strip is in global memory: the halo exchange may read it
				TRACE(1, q);
				if(!ended) { if(lag) { if(q>0u) tma_wait_read1(); } else tma_wait_read0(); } // the stage may be refilled once TMA has read it
				TRACE(2, q);
			}
			__syncwarp();
			if(!ended&&(!lag||q>0u)) issue_loads();
			if(leader) TRACE(3, q);
			if(park&&last_of_strip&&leader) { tma_wait_all_but(tiles_x-1u); mbar_arrive(bar_head); } // the strip's first tile (tiles_x-1 groups ago) is in global memory: its column 0 may be overwritten
			sxt = last_of_strip ? 0u : sxt+1u;
		}
		if(leader) tma_wait_all0();
		return;
	}

	// ---------------------------------------------------------------------------------------- consumers: two cells per thread
	const uint32_t row = tid/(uint32_t)(TX/2), lx = 2u*(tid%(uint32_t)(TX/2)), ly = row%(uint32_t)TY, lz = row/(uint32_t)TY;
	const float scale = (FAST&&P==P_FP16S) ? 32768.0f : 1.0f, inv = (FAST&&P==P_FP16S) ? 3.0517578E-5f : 1.0f;
	const uint64_t rowN = c.Px, planeN = (uint64_t)c.Px*c.Ny;
	const bool has_zones = VF&&(c.features&(F_NUDGING|F_SPONGE))!=0u;
	const int Nb = (c.features&F_NUDGING) ? (int)c.buffer_N : -1, Ns = (c.features&F_SPONGE) ? (int)c.sponge_N : 0;
	const uint32_t last_tx = c.Nx-(tiles_x-1u)*(uint32_t)TX; // cells of the last tile's rows that lie inside the lattice
	const uint32_t rowend_last = (last_tx-1u)&~1u; // local x of the pair that holds a row's last cell in the strip's last tile (odd Nx: that cell is the pair's cell 0)
	// walk: strips as published by the producer; inside a strip xt = 0..tiles_x-1; ring slot s and its phase advance with every tile
	uint32_t xt = 0u, s = 0u, ph = 0u, kstrip = 0u;
	int y0 = 0, z0 = 0, py0 = 0, pz0 = 0;
	uint32_t hb = 0u, phb = 0u; // this / the previous strip is a boundary strip of an overlapped halo exchange
	bool zone_yz = false; // this thread's row lies in a relaxation zone through its y / z position (decided once per strip)
	const int zone_xe = Nb>=0 ? (int)c.Nxg-1-Nb-63-c.Ox : 0x7FFFFFFF; // a warp (64 x-consecutive cells from local x = xw) reaches the east shell iff xw >= zone_xe, the west shell iff xw + Ox <= Nb
	for(uint32_t q=0u; ; q++) {
		const uint32_t s1 = s+1u==(uint32_t)S ? 0u : s+1u, ph1 = s+1u==(uint32_t)S ? ph^1u : ph;
		const int x0 = (int)xt*TX;
		uint8_t* const st = stage0+(size_t)s*CFG::STAGE_BYTES;
		uint8_t* const st1 = stage0+(size_t)s1*CFG::STAGE_BYTES;
		const bool first = xt==0u, last = xt+1u==tiles_x; // the next tile of the strip holds the +x slots of this tile's last column
		if(first) { // a new strip: which one?
			mbar_wait(bar_full+s, ph);
			const uint32_t raw = tile_strip[s];
			if(raw==END) break;
			const uint32_t strip = raw&~STRIP_BND;
			hb = raw&STRIP_BND;
			y0 = (int)(strip%tiles_y)*TY; z0 = (int)(strip/tiles_y)*TZ;
			if(has_zones) { const int yg = y0+(int)ly+c.Oy, zg = z0+(int)lz+c.Oz; zone_yz = yg<=Nb||yg>=(int)c.Nyg-1-Nb||zg>=(int)c.Nzg-1-Nb||zg>=(int)c.Nzg-2-Ns; }
		}
		const bool bnd_yz = y0==0||y0+TY>=(int)c.Ny||z0==0||z0+TZ>=(int)c.Nz; // strip touches the y/z boundary (uniform)
		const bool edge = bnd_yz||first||last;
		if(first&&bnd_yz) patch_yz<CFG, true>(c, st, x0, y0, z0, odd, tid, false);
		if(!last) { mbar_wait(bar_full+s1, ph1); if(bnd_yz) patch_yz<CFG, true>(c, st1, x0+TX, y0, z0, odd, tid, false); }
		if(bnd_yz) consumer_bar((uint32_t)NC);
		if(park&&kstrip>0u&&xt==1u) { // write the previous strip's periodic-x column
			mbar_wait(bar_head, (kstrip-1u)&1u);
			__syncwarp();
			if(lx==rowend_last) { flush_wrap<CFG>(c, stage0, (kstrip-1u)&1u, row, py0, pz0, odd); if(phb!=0u) { __threadfence(); atomicAdd(c.bdone, 1u); } } // the thread that wrote the parked elements last; one count per row of a boundary strip
		}

		if(tid==0u) TRACE(4, q);
		const uint32_t x = (uint32_t)x0+lx, y = (uint32_t)y0+ly, z = (uint32_t)z0+lz;
		const uint64_t n = (uint64_t)x+(uint64_t)y*rowN+(uint64_t)z*planeN;
		const uint32_t fl2 = ((const uint16_t*)(st+CFG::FLAG_OFF))[tid];
		const uint32_t fl0 = fl2&0xFFu, fl1 = fl2>>8;
		bool run0 = !((fl0&TYPE_BO)==TYPE_S||(fl0&TYPE_SU)==TYPE_G), run1 = !((fl1&TYPE_BO)==TYPE_S||(fl1&TYPE_SU)==TYPE_G);
		if(edge) {
			const bool in_yz = y<c.Ny&&z<c.Nz&&!((c.Dy>1u&&(y==0u||y>=c.Ny-1u))||(c.Dz>1u&&(z==0u||z>=c.Nz-1u)));
			run0 = run0&&in_yz&&x<c.Nx&&!(c.Dx>1u&&(x==0u||x>=c.Nx-1u));
			run1 = run1&&in_yz&&x+1u<c.Nx&&!(c.Dx>1u&&(x+1u>=c.Nx-1u));
		}
		if(EQ) { // rho/u of TYPE_E cells: into L2 one tile ahead
			if(!last) { // (the next strip's west face: prefetch_west_face, by the producer)
				const uint32_t fn = ((const uint16_t*)(st1+CFG::FLAG_OFF))[tid];
				if((fn&0x0003u)==TYPE_E||(fn&0x0300u)==(TYPE_E<<8)) { const uint64_t m = n+(uint64_t)TX; prefetch_l2(c.rho+m); prefetch_l2(c.u+m); prefetch_l2(c.u+c.N+m); prefetch_l2(c.u+2ull*c.N+m); }
			}
		}
		R* const box = (R*)st+tid; // the pair's word in box b is at byte offset box_off(b)
		const uint32_t par = kstrip&1u;
		uint8_t* const park_row = stage0+CFG::BOX_BYTES+(par*(uint32_t)CFG::ROWS+row)*(uint32_t)CFG::ES; // + box_off(b): this row's parked element of shifted box b
		if(park&&first&&lx==0u) { // park column 0 of the x-shifted boxes (pre-collision values; only the strip's last tile touches them)
#pragma unroll
			for(int b=0; b<Q; b++) if(box_shifted(b)) *(E*)(park_row+CFG::box_off(b)) = *(const E*)((const uint8_t*)box+CFG::box_off(b));
		}
		// the row-end lane (possibly in another warp) reads what the row's first lane parked. (It does so in the strip's last tile, and the stage ring keeps the warps
		// within STAGES tiles of each other, so for strips longer than the ring the barrier is not strictly needed; dropping it gained nothing: 49.2 vs 50.5 GLUP/s.)
		if(park&&first) { if(TX==64) __syncwarp(); else consumer_bar((uint32_t)NC); }
		if(run0||run1) {
			// element right of the pair's word in an x-shifted box: the next word of the row, or column 0 of the same row in the next stage / the parked column
			const uint32_t rowend_lx = last ? rowend_last : (uint32_t)(TX-2);
			uint8_t* nxt = (uint8_t*)(box+1);
			if(lx==rowend_lx) nxt = !last ? st1+(size_t)row*TX*CFG::ES : park ? park_row : wrap_x ? st+(size_t)row*TX*CFG::ES : (uint8_t*)box;
			const bool odd_end = last&&(c.Nx&1u)!=0u&&lx==rowend_lx; // odd Nx: cell 0 of the row's last pair is the lattice's last column, its +x element is `nxt` (cell 1 does not exist: run1 is false)
			const uint32_t bo0 = fl0&TYPE_BO, bo1 = fl1&TYPE_BO;
			const bool e0 = EQ&&run0&&bo0==TYPE_E, e1 = EQ&&run1&&bo1==TYPE_E;
			PairIn in;
			in.zones = false;
			if(has_zones) { // warp-uniform pre-test (a warp holds 64 x-consecutive cells of one row): does it reach into a relaxation zone? then gather the per-cell zone data now
				const int xw = x0+(int)(lx&~63u);
				if(zone_yz||xw+c.Ox<=Nb||xw>=zone_xe) {
					in.nudge_vertical = c.nudge_vertical;
					in.zr0 = zone_prefetch(c, x, y, z, run0&&bo0!=TYPE_E);
					in.zr1 = zone_prefetch(c, x+1u, y, z, run1&&bo1!=TYPE_E);
					in.zones = in.zr0.nudge||in.zr0.sponge||in.zr1.nudge||in.zr1.sponge;
				}
			}
			const auto store_fields = [&](const PairOut& o) { // rho / u of the non-TYPE_E cells (UPDATE_FIELDS, FX/kernel.cpp:1709-1715)
				const bool w0 = run0&&!e0, w1 = run1&&!e1;
				if(UF) {
					if(w0&&w1) {
						*(float2*)(c.rho+n) = o.rho.v; *(float2*)(c.u+n) = o.ux.v; *(float2*)(c.u+c.N+n) = o.uy.v; *(float2*)(c.u+2ull*c.N+n) = o.uz.v;
					} else {
						if(w0) { c.rho[n] = o.rho.v.x; c.u[n] = o.ux.v.x; c.u[c.N+n] = o.uy.v.x; c.u[2ull*c.N+n] = o.uz.v.x; }
						if(w1) { c.rho[n+1ull] = o.rho.v.y; c.u[n+1ull] = o.ux.v.y; c.u[c.N+n+1ull] = o.uy.v.y; c.u[2ull*c.N+n+1ull] = o.uz.v.y; }
					}
				}
				if(TH) store_upre(c, n, w0, w1, o); // thermal step: the velocity before the force half-step, for k_thermal_g
			};
			PairOut out;
			if constexpr (FAST&&CFG::TWOPASS) { // ---- two passes over the shared-memory boxes: moments, then relax + store (see lbm_vec.cuh)
				constexpr bool SG = (FEAT&F_SUBGRID)!=0u;
				uint8_t* const bb = (uint8_t*)box;
				const auto dec_in = [](const R w) -> f2 { if constexpr (P==P_FP16S) return PairCodec<P_FP16S>::dec_raw(w); else return PC::dec(w); };
				const auto enc_out = [](const f2 v) -> R { if constexpr (P==P_FP16S) return PairCodec<P_FP16S>::enc_raw(v); else if constexpr (P==P_FP16C) return PairCodec<P_FP16C>::enc_fast(v); else return PC::enc(v); };
				in.e0 = e0; in.e1 = e1; in.any_e = e0||e1; in.n = n; // TYPE_E lanes: rho / u are the boundary fields' (loaded in fast_prepare; prefetched into L2 a tile ago)
				Moments M;
				const f2 g0 = dec_in(*(const R*)bb);
				const auto ld1 = [&](const int k, f2& gi, f2& gj) {
					const R wa = *(const R*)(bb+CFG::box_off(1+2*k));
					R wb;
					if(pair_shifted(k)) wb = odd_end ? PC::shift_in_odd(nxt+CFG::box_off(2+2*k)) : PC::shift_in_e(bb+CFG::box_off(2+2*k), nxt+CFG::box_off(2+2*k));
					else wb = *(const R*)(bb+CFG::box_off(2+2*k));
					gi = dec_in(wa); gj = dec_in(wb);
				};
				moments_of<SG>(g0, ld1, M);
				FastK K;
				fast_prepare<FEAT>(c, a, in, M, scale, inv, K, out);
				store_fields(out); // now: rho / u need not stay live through pass 2
				asm volatile("" ::: "memory"); // pass 2 re-reads the boxes: the DDFs must not stay in registers
				const typename PC::M msk = PC::mask(run0, run1);
				*(R*)bb = PC::mixm(msk, enc_out(fma2(K.omw, g0, K.g0add)), *(const R*)bb);
				struct Raw { R wa, wb0; };
				const auto ld2 = [&](const int k, Raw& r, f2& gi, f2& gj) {
					r.wa = *(const R*)(bb+CFG::box_off(1+2*k));
					R wb;
					if(pair_shifted(k)) { wb = odd_end ? PC::shift_in_odd(nxt+CFG::box_off(2+2*k)) : PC::shift_in_e(bb+CFG::box_off(2+2*k), nxt+CFG::box_off(2+2*k)); r.wb0 = wb; }
					else wb = r.wb0 = *(const R*)(bb+CFG::box_off(2+2*k));
					gi = dec_in(r.wa); gj = dec_in(wb);
				};
				const auto st2 = [&](const int k, const Raw& r, const f2 gi, const f2 gj) { // f_i' goes to slot B, f_i+1' to slot A
					const int bA = 1+2*k, bB = 2+2*k;
					const R ni = enc_out(gi), nj = enc_out(gj);
					*(R*)(bb+CFG::box_off(bA)) = PC::mixm(msk, nj, r.wa);
					if(pair_shifted(k)) { if(odd_end) { if(run0) PC::shift_out_odd(nxt+CFG::box_off(bB), ni); } else PC::shift_out(bb+CFG::box_off(bB), nxt+CFG::box_off(bB), ni, run0, run1); }
					else *(R*)(bb+CFG::box_off(bB)) = PC::mixm(msk, ni, r.wb0);
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
			} else {
			// load_f: box 0 -> f0, box 1+2k -> f_(2k+1), box 2+2k -> f_(2k+2)
			f2 f[Q];
#pragma unroll
			for(int b=0; b<Q; b++) {
				R w;
				if(box_shifted(b)) w = odd_end ? PC::shift_in_odd(nxt+CFG::box_off(b)) : PC::shift_in_e((const uint8_t*)box+CFG::box_off(b), nxt+CFG::box_off(b));
				else w = *(const R*)((const uint8_t*)box+CFG::box_off(b));
				if(FAST&&P==P_FP16S) f[b] = PairCodec<P_FP16S>::dec_raw(*(const uint32_t*)&w); else f[b] = PC::dec(w);
			}
			float4 eb0 = make_float4(1.0f, 0.0f, 0.0f, 0.0f), eb1 = eb0; // boundary rho/u of TYPE_E lanes (prefetched into L2 one tile ago)
			if(FAST) {
				if(EQ&&__any_sync(__activemask(), e0||e1)) { // warp-uniform: the common path carries no TYPE_E code
					if(e0) eb0 = make_float4(c.rho[n], c.u[n], c.u[c.N+n], c.u[2ull*c.N+n]);
					if(e1) eb1 = make_float4(c.rho[n+1ull], c.u[n+1ull], c.u[c.N+n+1ull], c.u[2ull*c.N+n+1ull]);
					in.e0 = e0; in.e1 = e1;
					in.rho_e = mk2(eb0.x, eb1.x); in.ux_e = mk2(eb0.y, eb1.y); in.uy_e = mk2(eb0.z, eb1.z); in.uz_e = mk2(eb0.w, eb1.w);
					collide_fast2<FEAT, true>(c, a, in, f, scale, inv, out);
				} else {
					in.e0 = false; in.e1 = false;
					collide_fast2<FEAT, false>(c, a, in, f, scale, inv, out);
				}
			} else {
				if(e0) eb0 = make_float4(c.rho[n], c.u[n], c.u[c.N+n], c.u[2ull*c.N+n]);
				if(e1) eb1 = make_float4(c.rho[n+1ull], c.u[n+1ull], c.u[c.N+n+1ull], c.u[2ull*c.N+n+1ull]);
				collide_strict2<FEAT>(c, a, in, f, out);
			}
			if(!FAST&&(e0||e1)) { // TYPE_E lanes, bit-exact path: f := feq(rho, u) of the boundary fields in scalar code
#pragma unroll 1
				for(uint32_t l=0u; l<2u; l++) {
					if(!(l==0u ? e0 : e1)) continue;
					float feq[Q];
					const float4 eb = l==0u ? eb0 : eb1;
					equilibrium_cell<FEAT>(c, a, eb.x, eb.y, eb.z, eb.w, scale, feq);
#pragma unroll
					for(int b=0; b<Q; b++) { if(l==0u) f[b].v.x = feq[b]; else f[b].v.y = feq[b]; }
				}
			}
			// store_f: f_i' goes to slot B (box 2+2k), f_i+1' to slot A (box 1+2k)
			R nw[Q];
#pragma unroll
			for(int b=0; b<Q; b++) {
				if(FAST&&P==P_FP16S) { const uint32_t r_ = PairCodec<P_FP16S>::enc_raw(f[b]); nw[b] = *(const R*)&r_; }
				else if constexpr (FAST&&P==P_FP16C) nw[b] = PairCodec<P_FP16C>::enc_fast(f[b]);
				else nw[b] = PC::enc(f[b]);
			}
			uint8_t* const bb = (uint8_t*)box;
			if(run0&&run1) {
				*(R*)bb = nw[0];
#pragma unroll
				for(int k=0; k<9; k++) {
					const int bA = 1+2*k, bB = 2+2*k;
					*(R*)(bb+CFG::box_off(bA)) = nw[bB];
					if(box_shifted(bB)) PC::shift_out_both(bb+CFG::box_off(bB), nxt+CFG::box_off(bB), nw[bA]);
					else *(R*)(bb+CFG::box_off(bB)) = nw[bA];
				}
			} else {
				*(R*)bb = PC::mix(run0, run1, nw[0], *(R*)bb);
#pragma unroll
				for(int k=0; k<9; k++) {
					const int bA = 1+2*k, bB = 2+2*k;
					*(R*)(bb+CFG::box_off(bA)) = PC::mix(run0, run1, nw[bB], *(R*)(bb+CFG::box_off(bA)));
					if(box_shifted(bB)) { if(odd_end) { if(run0) PC::shift_out_odd(nxt+CFG::box_off(bB), nw[bA]); } else PC::shift_out(bb+CFG::box_off(bB), nxt+CFG::box_off(bB), nw[bA], run0, run1); }
					else *(R*)(bb+CFG::box_off(bB)) = PC::mix(run0, run1, nw[bA], *(R*)(bb+CFG::box_off(bB)));
				}
			}
			store_fields(out);
			}
		}

		if(bnd_yz) { // write the wrapped elements back to their real addresses (TMA clips them)
			consumer_bar((uint32_t)NC);
			patch_yz<CFG, false>(c, st, x0, y0, z0, odd, tid, park);
		}
		fence_async_smem(); // make this thread's shared-memory writes visible to the TMA store
		if(tid==0u) TRACE(5, q);
#ifndef LUW_PRODUCER_NAMED_BARRIER
		// One arrival per warp: every lane has fenced its own shared-memory writes for the async proxy, __syncwarp orders them before lane 0's releasing arrive.
		// (With one arrival per THREAD the producer's try_wait woke up ~94 times per tile -- once per arrival it could see -- and its polling was 15 % of all
		// issued instructions on the channel case, profiles/r1c_ncu_channel512_fp16s.md.)
		__syncwarp();
		if((tid&31u)==0u) mbar_arrive(bar_done+s);
#else
		bar_arrive_id(2u+s, (uint32_t)CFG::THREADS);
#endif
		s = s1; ph = ph1;
		if(last) { xt = 0u; py0 = y0; pz0 = z0; phb = hb; kstrip++; } else xt++;
	}
	TRACE_CLK(2);
	if(park&&kstrip>0u) { // the last strip's periodic-x column
		mbar_wait(bar_head, (kstrip-1u)&1u);
		if(lx==rowend_last) { flush_wrap<CFG>(c, stage0, (kstrip-1u)&1u, row, py0, pz0, odd); if(phb!=0u) { __threadfence(); atomicAdd(c.bdone, 1u); } }
	}
}

} // anonymous namespace
} // namespace luw
