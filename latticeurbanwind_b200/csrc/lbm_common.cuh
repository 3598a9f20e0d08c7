// Shared device-side definitions for the LBM step kernels (sm_100a).
// Layout and semantics follow the reference device code FX/kernel.cpp (FX = core/cfd_core/FluidX3D/src); every helper cites
// the lines it re-implements. Nothing here is a translation of the OpenCL text: indexing is 3-D without div/mod, the
// per-case constants are run-time parameters, and relaxation-zone weights come from host-built tables.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace luw {

constexpr int Q = 19;
constexpr uint32_t TYPE_S = 0x01u, TYPE_E = 0x02u, TYPE_BO = 0x03u, TYPE_G = 0x20u, TYPE_SU = 0x38u; // FX/defines.hpp:50-57, FX/lbm.cpp:689-694
constexpr uint32_t TYPE_T = 0x04u; // temperature boundary, FX/defines.hpp:52
constexpr float LAT_C = 0.57735027f; // def_c, FX/lbm.cpp:662
constexpr float W0 = 1.0f/3.0f, WS = 1.0f/18.0f, WE = 1.0f/36.0f; // FX/lbm.cpp:672-674

enum : uint32_t { F_UPDATE_FIELDS = 1u, F_VOLUME_FORCE = 2u, F_EQUILIBRIUM = 4u, F_SUBGRID = 8u, F_NUDGING = 16u, F_SPONGE = 32u, F_TEMPERATURE = 64u };
enum : int { P_FP32 = 0, P_FP16S = 1, P_FP16C = 2 };

struct DomainConst { // the def_* constants of FX/lbm.cpp:612-783 as kernel parameters
	uint32_t Nx, Ny, Nz;
	uint32_t Px; // row pitch in elements: Nx rounded up to a multiple of 16, so that every row of every array starts 16-byte aligned (TMA); host images stay dense
	uint64_t N; // elements per component / DDF slot ON THE DEVICE: Px*Ny*Nz (the device index of cell (x,y,z) is x+(y+z*Ny)*Px)
	uint32_t Dx, Dy, Dz;
	int Ox, Oy, Oz;
	uint32_t Nxg, Nyg, Nzg; // global lattice
	int wx, ex, sy, ny, tz; // local index of the global west/east/south/north/top boundary plane
	int has_w, has_e, has_s, has_n, has_t;
	float w; // def_w
	float tau0, tau0sq; // 1/def_w and its square (IEEE single, as the Smagorinsky term forms them: FX/kernel.cpp:1735)
	int precision; // P_FP32 / P_FP16S / P_FP16C
	uint32_t features;
	int downstream_face;
	uint32_t buffer_N; float buffer_inv_tau; int nudge_vertical;
	uint32_t sponge_N;
	const float* wbuf; // [buffer_N+1] sin^2 ramp by distance, built on the host (FX/kernel.cpp:1579-1581)
	const float* sigma; // [sponge_N] inv_tau*sin^2 ramp by depth (FX/kernel.cpp:1603-1605)
	void* fi; float* rho; float* u; uint8_t* flags;
	uint32_t* sched; // strip counter of the tiled step (zeroed in-stream before every launch)
	// thermal D3Q7 extension (F_TEMPERATURE; FX/lbm.cpp:322-323, 750-752): 7 x N DDFs in the storage type of fi, N floats, def_w_T / def_beta / def_T_avg
	void* gi; float* T;
	float w_T, beta, T_avg;
	// Strip order of the tiled step when the halo exchange overlaps it (luw_step_halo_ipc; 0 everywhere = natural order, nothing signalled): the strips that hold the
	// halo / boundary layers of the decomposed y and z axes (so_ylo / so_yhi tile rows at the low / high end of y, so_zlo / so_zhi tile planes in z; so_nb strips in
	// all) are handed out first, and every finished boundary strip is counted in *bdone once its DDFs are in global memory.
	uint32_t so_ylo, so_yhi, so_zlo, so_zhi, so_nb;
	uint32_t* bdone;
	// Thermal step in two kernels (luw_cabi.cu enqueue_step): the TMA-tiled momentum kernel leaves the velocity BEFORE the force half-step here (3 x N floats, the
	// one input of the g collision that is never stored otherwise, FX/kernel.cpp:1664), k_thermal_g consumes it. nullptr on domains without TEMPERATURE.
	float* upre;
};
struct StepArgs { uint64_t t; float fx, fy, fz, ox, oy, oz; }; // per-step kernel arguments, FX/lbm.cpp:345
// Strip order of an overlapped step for tiles of TY x TZ rows / planes: the tile rows / planes that hold layers 0, 1, N-2, N-1 of the decomposed y / z axes come first.
// false (so_nb = 0: natural order) when nothing would be left to overlap with.
inline bool strip_order_fill(DomainConst& o, const uint32_t TY, const uint32_t TZ) {
	const uint32_t Ty = (o.Ny+TY-1u)/TY, Tz = (o.Nz+TZ-1u)/TZ;
	o.so_ylo = o.Dy>1u ? 1u/TY+1u : 0u; o.so_yhi = o.Dy>1u ? Ty-(o.Ny-2u)/TY : 0u;
	o.so_zlo = o.Dz>1u ? 1u/TZ+1u : 0u; o.so_zhi = o.Dz>1u ? Tz-(o.Nz-2u)/TZ : 0u;
	o.so_nb = 0u;
	if(o.so_ylo+o.so_yhi>=Ty||o.so_zlo+o.so_zhi>=Tz) return false;
	o.so_nb = (o.so_zlo+o.so_zhi)*Ty+(o.so_ylo+o.so_yhi)*(Tz-o.so_zlo-o.so_zhi);
	return o.so_nb>0u;
}
// i-th strip handed out by the counter -> strip id ty + tz*Ty: the so_nb boundary strips first (whole tile planes at the z ends, then the y-end tile rows of the inner planes), then the interior
__host__ __device__ inline uint32_t strip_of(const DomainConst& c, const uint32_t i, const uint32_t Ty, const uint32_t Tz) {
	if(c.so_nb==0u) return i;
	const uint32_t Bz = (c.so_zlo+c.so_zhi)*Ty, per = c.so_ylo+c.so_yhi;
	uint32_t ty, tz;
	if(i<Bz) { const uint32_t k = i/Ty; tz = k<c.so_zlo ? k : Tz-c.so_zhi+(k-c.so_zlo); ty = i%Ty; }
	else if(i<c.so_nb) { const uint32_t j = i-Bz, k = j%per; tz = c.so_zlo+j/per; ty = k<c.so_ylo ? k : Ty-c.so_yhi+(k-c.so_ylo); }
	else { const uint32_t j = i-c.so_nb, Iy = Ty-per; ty = c.so_ylo+j%Iy; tz = c.so_zlo+j/Iy; }
	return ty+tz*Ty;
}

namespace { // internal linkage: this header is compiled into two translation units with different arithmetic flags

// ------------------------------------------------------------------ DDF codecs (FX/lbm.cpp:706-721, FX/kernel.cpp:864-875)
template<int P> struct Ddf;
template<> struct Ddf<P_FP32> {
	typedef float T;
	static __device__ __forceinline__ float dec(const float x) { return x; }
	static __device__ __forceinline__ float enc(const float x) { return x; }
};
template<> struct Ddf<P_FP16S> { // IEEE half holding 2^15 * f
	typedef uint16_t T;
	static __device__ __forceinline__ float dec(const uint16_t x) { return __half2float(__ushort_as_half(x))*3.0517578E-5f; }
	static __device__ __forceinline__ uint16_t enc(const float x) { return __half_as_ushort(__float2half_rn(x*32768.0f)); }
};
template<> struct Ddf<P_FP16C> { // custom 1-4-11 format, bias 15, range +-2, with subnormals
	typedef uint16_t T;
	// decode (FX/kernel.cpp:864-869) without branches: the 15 magnitude bits, shifted into a float's exponent/mantissa fields, read as 2^-112 times the
	// value -- for the subnormal codes (e = 0) too, because the float is then itself subnormal -- and one exact multiplication by 2^112 restores it.
	static __device__ __forceinline__ float dec(const uint16_t h) {
		const uint32_t x = h;
		return __uint_as_float(__float_as_uint(__fmul_rn(__uint_as_float((x&0x7FFFu)<<12), 5.192296858534828e33f))|((x&0x8000u)<<16));
	}
	static __device__ __forceinline__ uint16_t enc(const float f) {
		const uint32_t b = __float_as_uint(f)+0x00000800u, e = (b&0x7F800000u)>>23, m = b&0x007FFFFFu;
		uint32_t r = (b&0x80000000u)>>16;
		if(e>112u) r |= (((e-112u)<<11)&0x7800u)|(m>>12);
		if(e<113u&&e>100u) r |= (((0x007FF800u+m)>>(124u-e))+1u)>>1;
		return (uint16_t)r;
	}
};

// ------------------------------------------------------------------ moments, equilibrium, forcing (FX/kernel.cpp:1016-1113)
__device__ __forceinline__ void f_eq(const float rho, float ux, float uy, float uz, float* feq) {
	const float rhom1 = rho-1.0f;
	const float c3 = -3.0f*(ux*ux+uy*uy+uz*uz);
	uz *= 3.0f; ux *= 3.0f; uy *= 3.0f;
	feq[0] = W0*fmaf(rho, 0.5f*c3, rhom1);
	const float u0 = ux+uy, u1 = ux+uz, u2 = uy+uz, u3 = ux-uy, u4 = ux-uz, u5 = uy-uz;
	const float rhos = WS*rho, rhoe = WE*rho, rhom1s = WS*rhom1, rhom1e = WE*rhom1;
#define LUW_FEQ_PAIR(i, a, r, r1) { const float q_ = fmaf(a, a, c3); feq[i] = fmaf(r, fmaf(0.5f, q_, a), r1); feq[i+1] = fmaf(r, fmaf(0.5f, q_, -(a)), r1); }
	LUW_FEQ_PAIR( 1, ux, rhos, rhom1s) LUW_FEQ_PAIR( 3, uy, rhos, rhom1s) LUW_FEQ_PAIR( 5, uz, rhos, rhom1s)
	LUW_FEQ_PAIR( 7, u0, rhoe, rhom1e) LUW_FEQ_PAIR( 9, u1, rhoe, rhom1e) LUW_FEQ_PAIR(11, u2, rhoe, rhom1e)
	LUW_FEQ_PAIR(13, u3, rhoe, rhom1e) LUW_FEQ_PAIR(15, u4, rhoe, rhom1e) LUW_FEQ_PAIR(17, u5, rhoe, rhom1e)
#undef LUW_FEQ_PAIR
}
__device__ __forceinline__ void rho_u(const float* f, float& rho, float& ux, float& uy, float& uz) {
	float r = f[0];
#pragma unroll
	for(int i=1; i<Q; i++) r += f[i];
	r += 1.0f;
	const float mx = f[ 1]-f[ 2]+f[ 7]-f[ 8]+f[ 9]-f[10]+f[13]-f[14]+f[15]-f[16];
	const float my = f[ 3]-f[ 4]+f[ 7]-f[ 8]+f[11]-f[12]+f[14]-f[13]+f[17]-f[18];
	const float mz = f[ 5]-f[ 6]+f[ 9]-f[10]+f[11]-f[12]+f[16]-f[15]+f[18]-f[17];
	rho = r; ux = mx/r; uy = my/r; uz = mz/r;
}
// Guo forcing terms; the c_i are written out (0/+-1 factors of the reference's generic loop are exact and dropped)
__device__ __forceinline__ void forcing_terms(const float ux, const float uy, const float uz, const float fx, const float fy, const float fz, float* Fin) {
	const float uF = -0.33333334f*fmaf(ux, fx, fmaf(uy, fy, uz*fz));
	const float t3 = 0.33333334f;
	Fin[0] = 9.0f*W0*uF;
	const float ks = 9.0f*WS, ke = 9.0f*WE;
#define LUW_FIN_PAIR(i, k, cf, cu) { const float cf_ = (cf), cu_ = (cu); Fin[i] = k*fmaf(cf_, cu_+t3, uF); Fin[i+1] = k*fmaf(-cf_, -cu_+t3, uF); }
	LUW_FIN_PAIR( 1, ks, fx, ux) LUW_FIN_PAIR( 3, ks, fy, uy) LUW_FIN_PAIR( 5, ks, fz, uz)
	LUW_FIN_PAIR( 7, ke, fx+fy, ux+uy) LUW_FIN_PAIR( 9, ke, fx+fz, ux+uz) LUW_FIN_PAIR(11, ke, fy+fz, uy+uz)
	LUW_FIN_PAIR(13, ke, fx-fy, ux-uy) LUW_FIN_PAIR(15, ke, fx-fz, ux-uz) LUW_FIN_PAIR(17, ke, fy-fz, uy-uz)
#undef LUW_FIN_PAIR
}
__device__ __forceinline__ float clampc(const float x) { return fminf(fmaxf(x, -LAT_C), LAT_C); }

// Smagorinsky-Lilly relaxation rate, FX/kernel.cpp:1723-1736 (zero-weight terms of the reference's generic loop dropped, order kept)
__device__ __forceinline__ float smagorinsky_w(const float w0, const float* f, const float* feq, const float rho) {
	float n[Q];
#pragma unroll
	for(int i=1; i<Q; i++) n[i] = f[i]-feq[i];
	const float Hxx = n[1]+n[2]+n[7]+n[8]+n[9]+n[10]+n[13]+n[14]+n[15]+n[16];
	const float Hyy = n[3]+n[4]+n[7]+n[8]+n[11]+n[12]+n[13]+n[14]+n[17]+n[18];
	const float Hzz = n[5]+n[6]+n[9]+n[10]+n[11]+n[12]+n[15]+n[16]+n[17]+n[18];
	const float Hxy = n[7]+n[8]-n[13]-n[14];
	const float Hxz = n[9]+n[10]-n[15]-n[16];
	const float Hyz = n[11]+n[12]-n[17]-n[18];
	const float tau0 = 1.0f/w0;
	const float Qn = Hxx*Hxx+Hyy*Hyy+Hzz*Hzz+2.0f*(Hxy*Hxy+Hxz*Hxz+Hyz*Hyz);
	return 2.0f/(tau0+sqrtf(tau0*tau0+0.76421222f*sqrtf(Qn)/rho));
}

// ------------------------------------------------------------------ body force of the LUW step (FX/kernel.cpp:1516-1614)
// Coriolis always; buffer nudging and top sponge for non-TYPE_E cells inside the relaxation zones this domain owns.
__device__ __forceinline__ void luw_force(const DomainConst& c, const StepArgs& a, const uint32_t x, const uint32_t y, const uint32_t z, const uint32_t bo,
	const bool zones, const float rho, const float ux, const float uy, const float uz, float& Fx, float& Fy, float& Fz) {
	float fxn = a.fx, fyn = a.fy, fzn = a.fz;
	fxn += -2.0f*rho*(a.oy*uz-a.oz*uy);
	fyn += -2.0f*rho*(a.oz*ux-a.ox*uz);
	fzn += -2.0f*rho*(a.ox*uy-a.oy*ux);
	if(zones&&bo!=TYPE_E) {
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
				const float k = __ldg(c.wbuf+dmin)*c.buffer_inv_tau;
				const float ax = k*(__ldg(c.u+nref)-ux);
				const float ay = k*(__ldg(c.u+c.N+nref)-uy);
				const float az = c.nudge_vertical==1 ? k*(__ldg(c.u+2ull*c.N+nref)-uz) : 0.0f;
				fxn += rho*ax; fyn += rho*ay; fzn += rho*az;
			}
		}
		if((c.features&F_SPONGE)&&c.has_t) {
			const int dt = (int)(c.Nzg-2u)-((int)z+c.Oz);
			if(dt>=0&&dt<(int)c.sponge_N) {
				const float s = __ldg(c.sigma+dt);
				const uint64_t nref = x+y*row+(uint64_t)c.tz*plane;
				fxn += rho*s*(__ldg(c.u+nref)-ux);
				fyn += rho*s*(__ldg(c.u+c.N+nref)-uy);
				fzn += rho*s*(__ldg(c.u+2ull*c.N+nref)-uz);
			}
		}
	}
	Fx = fxn; Fy = fyn; Fz = fzn;
}

__device__ __forceinline__ bool is_halo(const DomainConst& c, const uint32_t x, const uint32_t y, const uint32_t z) { // FX/kernel.cpp:856-859
	return (c.Dx>1u&&(x==0u||x>=c.Nx-1u))||(c.Dy>1u&&(y==0u||y>=c.Ny-1u))||(c.Dz>1u&&(z==0u||z>=c.Nz-1u));
}

// linear offsets of the 9 "+" neighbours n+c_i (i odd) with periodic wrap, FX/kernel.cpp:920-958
struct Nbr { uint64_t j[Q]; };
__device__ __forceinline__ void neighbors(const DomainConst& c, const uint32_t x, const uint32_t y, const uint32_t z, uint64_t* j) {
	const uint64_t row = c.Px, plane = (uint64_t)c.Px*c.Ny;
	const uint64_t x0 = x, xp = x+1u==c.Nx ? 0u : x+1u, xm = x==0u ? c.Nx-1u : x-1u;
	const uint64_t y0 = y*row, yp = (y+1u==c.Ny ? 0u : y+1u)*row, ym = (y==0u ? c.Ny-1u : y-1u)*row;
	const uint64_t z0 = z*plane, zp = (z+1u==c.Nz ? 0u : z+1u)*plane, zm = (z==0u ? c.Nz-1u : z-1u)*plane;
	j[ 0] = x0+y0+z0;
	j[ 1] = xp+y0+z0; j[ 2] = xm+y0+z0; j[ 3] = x0+yp+z0; j[ 4] = x0+ym+z0; j[ 5] = x0+y0+zp; j[ 6] = x0+y0+zm;
	j[ 7] = xp+yp+z0; j[ 8] = xm+ym+z0; j[ 9] = xp+y0+zp; j[10] = xm+y0+zm; j[11] = x0+yp+zp; j[12] = x0+ym+zm;
	j[13] = xp+ym+z0; j[14] = xm+yp+z0; j[15] = xp+y0+zm; j[16] = xm+y0+zp; j[17] = x0+yp+zm; j[18] = x0+ym+zp;
}

} // anonymous namespace
} // namespace luw
