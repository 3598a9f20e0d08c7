/* TEST INFRASTRUCTURE -- see luw_oracle.h for scope, citations and the parity pin.
 * Build: gcc -std=c11 -O2 -ffp-contract=off -fopenmp -fPIC -shared luw_oracle.c -lm -o libluw_oracle.so
 * FX = /root/reference/core/cfd_core/FluidX3D/src
 */
#include "luw_oracle.h"
#include <math.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define Q 19
#define TYPE_S 0x01u
#define TYPE_E 0x02u
#define TYPE_BO 0x03u
#define TYPE_G 0x20u
#define TYPE_SU 0x38u
#define LAT_C 0.57735027f /* def_c, FX/lbm.cpp:662 */
#define W0 (1.0f/3.0f) /* FX/lbm.cpp:672-674 */
#define WS (1.0f/18.0f)
#define WE (1.0f/36.0f)

static int g_threads = 0;
void luwo_set_threads(int n) { g_threads = n; }
int luwo_get_threads(void) {
#ifdef _OPENMP
	return g_threads>0 ? g_threads : omp_get_max_threads();
#else
	return 1;
#endif
}
#ifdef _OPENMP
#define PAR_FOR _Pragma("omp parallel for schedule(static) num_threads(luwo_get_threads())")
#else
#define PAR_FOR
#endif

/* ------------------------------------------------------------------ bit casts and codecs */
static inline uint32_t f2u(float x) { uint32_t r; memcpy(&r, &x, 4); return r; }
static inline float u2f(uint32_t x) { float r; memcpy(&r, &x, 4); return r; }

float luwo_half_to_float(uint16_t h) { /* what OpenCL's vload_half does (FX/lbm.cpp:709) */
	const uint32_t s = (uint32_t)(h&0x8000u)<<16;
	uint32_t e = (h>>10)&0x1Fu, m = h&0x3FFu;
	if(e==0u) {
		if(m==0u) return u2f(s);
		int k = 0;
		while(!(m&0x400u)) { m <<= 1; k++; }
		return u2f(s|((uint32_t)(113-k)<<23)|((m&0x3FFu)<<13));
	}
	if(e==31u) return u2f(s|0x7F800000u|(m<<13));
	return u2f(s|((e+112u)<<23)|(m<<13));
}
uint16_t luwo_float_to_half_rte(float f) { /* what OpenCL's vstore_half_rte does (FX/lbm.cpp:710) */
	const uint32_t x = f2u(f), s = (x>>16)&0x8000u, a = x&0x7FFFFFFFu;
	if(a>0x7F800000u) return (uint16_t)(s|0x7E00u|((a>>13)&0x3FFu)); /* nan */
	if(a>=0x477FF000u) return (uint16_t)(s|0x7C00u); /* overflow (>=65520) and inf */
	if(a<=0x33000000u) return (uint16_t)s; /* |f|<=2^-25 -> 0 (tie goes to even = 0) */
	const int e = (int)(a>>23)-127;
	const uint32_t m = (a&0x007FFFFFu)|0x00800000u;
	const int shift = e<-14 ? 13+(-14-e) : 13;
	const uint32_t he = e<-14 ? 0u : (uint32_t)(e+15);
	const uint32_t keep = m>>shift, rem = m&((1u<<shift)-1u), half = 1u<<(shift-1);
	uint32_t r = (he==0u ? keep : (keep&0x3FFu))|(he<<10);
	if(rem>half||(rem==half&&(keep&1u))) r++;
	return (uint16_t)(s|r);
}
float luwo_fp16c_to_float(uint16_t x) { /* FX/kernel.cpp:864-869: 1-4-11 format, exponent bias 15 */
	const uint32_t e = ((uint32_t)x&0x7800u)>>11;
	const uint32_t m = ((uint32_t)x&0x07FFu)<<12;
	const uint32_t v = f2u((float)m)>>23; /* exponent of m as float = position of its leading one */
	uint32_t r = ((uint32_t)x&0x8000u)<<16;
	if(e!=0u) r |= ((e+112u)<<23)|m;
	else if(m!=0u) r |= ((v-37u)<<23)|((m<<(150u-v))&0x007FF000u);
	return u2f(r);
}
uint16_t luwo_float_to_fp16c(float x) { /* FX/kernel.cpp:870-875 (device version: no saturation term) */
	const uint32_t b = f2u(x)+0x00000800u;
	const uint32_t e = (b&0x7F800000u)>>23;
	const uint32_t m = b&0x007FFFFFu;
	uint32_t r = (b&0x80000000u)>>16;
	if(e>112u) r |= (((e-112u)<<11)&0x7800u)|(m>>12);
	if(e<113u&&e>100u) r |= (((0x007FF800u+m)>>(124u-e))+1u)>>1;
	return (uint16_t)r;
}

static inline float ddf_load(const luwo_params* p, const void* fi, uint64_t idx) {
	switch(p->precision) {
		case LUWO_FP16S: return luwo_half_to_float(((const uint16_t*)fi)[idx])*3.0517578E-5f;
		case LUWO_FP16C: return luwo_fp16c_to_float(((const uint16_t*)fi)[idx]);
		default: return ((const float*)fi)[idx];
	}
}
static inline void ddf_store(const luwo_params* p, void* fi, uint64_t idx, float x) {
	switch(p->precision) {
		case LUWO_FP16S: ((uint16_t*)fi)[idx] = luwo_float_to_half_rte(x*32768.0f); break;
		case LUWO_FP16C: ((uint16_t*)fi)[idx] = luwo_float_to_fp16c(x); break;
		default: ((float*)fi)[idx] = x;
	}
}

/* ------------------------------------------------------------------ lattice tables, FX/kernel.cpp:880-919 (D3Q19 rows) */
static const float CX[Q] = {0, 1,-1, 0, 0, 0, 0, 1,-1, 1,-1, 0, 0, 1,-1, 1,-1, 0, 0};
static const float CY[Q] = {0, 0, 0, 1,-1, 0, 0, 1,-1, 0, 0, 1,-1,-1, 1, 0, 0, 1,-1};
static const float CZ[Q] = {0, 0, 0, 0, 0, 1,-1, 0, 0, 1,-1, 1,-1, 0, 0,-1, 1,-1, 1};
static const float WT[Q] = {W0, WS,WS,WS,WS,WS,WS, WE,WE,WE,WE,WE,WE,WE,WE,WE,WE,WE,WE};

/* ------------------------------------------------------------------ indexing, FX/kernel.cpp:833-839, 856-859, 920-958 */
typedef struct { uint32_t x, y, z; } xyz_t;
static inline uint64_t cells(const luwo_params* p) { return (uint64_t)p->Nx*(uint64_t)p->Ny*(uint64_t)p->Nz; }
static inline xyz_t coords(const luwo_params* p, uint64_t n) {
	const uint64_t plane = (uint64_t)p->Nx*p->Ny;
	const uint32_t t = (uint32_t)(n%plane);
	xyz_t r = { t%p->Nx, t/p->Nx, (uint32_t)(n/plane) };
	return r;
}
static inline uint64_t lin(const luwo_params* p, uint32_t x, uint32_t y, uint32_t z) {
	return (uint64_t)x+((uint64_t)y+(uint64_t)z*p->Ny)*(uint64_t)p->Nx;
}
static inline int is_halo(const luwo_params* p, xyz_t c) {
	return (p->Dx>1u&&(c.x==0u||c.x>=p->Nx-1u))||(p->Dy>1u&&(c.y==0u||c.y>=p->Ny-1u))||(p->Dz>1u&&(c.z==0u||c.z>=p->Nz-1u));
}
static void neighbors(const luwo_params* p, uint64_t n, uint64_t* j) {
	const xyz_t c = coords(p, n);
	const uint64_t row = p->Nx, plane = (uint64_t)p->Nx*p->Ny;
	const uint64_t x0 = c.x, xp = (c.x+1u)%p->Nx, xm = (c.x+p->Nx-1u)%p->Nx;
	const uint64_t y0 = c.y*row, yp = ((c.y+1u)%p->Ny)*row, ym = ((c.y+p->Ny-1u)%p->Ny)*row;
	const uint64_t z0 = c.z*plane, zp = ((c.z+1u)%p->Nz)*plane, zm = ((c.z+p->Nz-1u)%p->Nz)*plane;
	j[ 0] = n;
	j[ 1] = xp+y0+z0; j[ 2] = xm+y0+z0;
	j[ 3] = x0+yp+z0; j[ 4] = x0+ym+z0;
	j[ 5] = x0+y0+zp; j[ 6] = x0+y0+zm;
	j[ 7] = xp+yp+z0; j[ 8] = xm+ym+z0;
	j[ 9] = xp+y0+zp; j[10] = xm+y0+zm;
	j[11] = x0+yp+zp; j[12] = x0+ym+zm;
	j[13] = xp+ym+z0; j[14] = xm+yp+z0;
	j[15] = xp+y0+zm; j[16] = xm+y0+zp;
	j[17] = x0+yp+zm; j[18] = x0+ym+zp;
}

/* ------------------------------------------------------------------ Esoteric-Pull, FX/kernel.cpp:1338-1351 */
static void load_f(const luwo_params* p, uint64_t n, float* f, const void* fi, const uint64_t* j, uint64_t t) {
	const uint64_t N = cells(p);
	const int odd = (int)(t&1u);
	f[0] = ddf_load(p, fi, n);
	for(uint32_t i=1u; i<Q; i+=2u) {
		f[i   ] = ddf_load(p, fi, (uint64_t)(odd ? i    : i+1u)*N+n   );
		f[i+1u] = ddf_load(p, fi, (uint64_t)(odd ? i+1u : i   )*N+j[i]);
	}
}
static void store_f(const luwo_params* p, uint64_t n, const float* f, void* fi, const uint64_t* j, uint64_t t) {
	const uint64_t N = cells(p);
	const int odd = (int)(t&1u);
	ddf_store(p, fi, n, f[0]);
	for(uint32_t i=1u; i<Q; i+=2u) {
		ddf_store(p, fi, (uint64_t)(odd ? i+1u : i   )*N+j[i], f[i   ]);
		ddf_store(p, fi, (uint64_t)(odd ? i    : i+1u)*N+n   , f[i+1u]);
	}
}

/* ------------------------------------------------------------------ moments and equilibrium, FX/kernel.cpp:1016-1100 */
void luwo_calculate_f_eq(float rho, float ux, float uy, float uz, float* feq) {
	const float rhom1 = rho-1.0f;
	const float c3 = -3.0f*(ux*ux+uy*uy+uz*uz);
	uz *= 3.0f; ux *= 3.0f; uy *= 3.0f;
	feq[0] = W0*fmaf(rho, 0.5f*c3, rhom1);
	const float u0=ux+uy, u1=ux+uz, u2=uy+uz, u3=ux-uy, u4=ux-uz, u5=uy-uz;
	const float rhos=WS*rho, rhoe=WE*rho, rhom1s=WS*rhom1, rhom1e=WE*rhom1;
	const float a[9] = {ux, uy, uz, u0, u1, u2, u3, u4, u5};
	for(int k=0; k<9; k++) {
		const float r = k<3 ? rhos : rhoe, r1 = k<3 ? rhom1s : rhom1e;
		const float q = fmaf(a[k], a[k], c3);
		feq[2*k+1] = fmaf(r, fmaf(0.5f, q,  a[k]), r1);
		feq[2*k+2] = fmaf(r, fmaf(0.5f, q, -a[k]), r1);
	}
}
static void rho_u(const float* f, float* rhon, float* uxn, float* uyn, float* uzn) {
	float rho = f[0];
	for(int i=1; i<Q; i++) rho += f[i];
	rho += 1.0f;
	const float ux = f[ 1]-f[ 2]+f[ 7]-f[ 8]+f[ 9]-f[10]+f[13]-f[14]+f[15]-f[16];
	const float uy = f[ 3]-f[ 4]+f[ 7]-f[ 8]+f[11]-f[12]+f[14]-f[13]+f[17]-f[18];
	const float uz = f[ 5]-f[ 6]+f[ 9]-f[10]+f[11]-f[12]+f[16]-f[15]+f[18]-f[17];
	*rhon = rho; *uxn = ux/rho; *uyn = uy/rho; *uzn = uz/rho;
}
static void forcing_terms(float ux, float uy, float uz, float fx, float fy, float fz, float* Fin) { /* FX/kernel.cpp:1103-1113 */
	const float uF = -0.33333334f*fmaf(ux, fx, fmaf(uy, fy, uz*fz));
	Fin[0] = 9.0f*W0*uF;
	for(int i=1; i<Q; i++) Fin[i] = 9.0f*WT[i]*fmaf(CX[i]*fx+CY[i]*fy+CZ[i]*fz, CX[i]*ux+CY[i]*uy+CZ[i]*uz+0.33333334f, uF);
}
static inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

/* global-frame helpers shared by nudging and sponge, FX/lbm.cpp:613-627 */
typedef struct { uint32_t Nxg, Nyg, Nzg; int wx, ex, sy, ny, tz; int has_w, has_e, has_s, has_n, has_t; } frame_t;
static frame_t frame(const luwo_params* p) {
	frame_t g;
	g.Nxg = (p->Nx-2u*(p->Dx>1u))*p->Dx; g.Nyg = (p->Ny-2u*(p->Dy>1u))*p->Dy; g.Nzg = (p->Nz-2u*(p->Dz>1u))*p->Dz;
	g.wx = -p->Ox; g.ex = (int)g.Nxg-1-p->Ox; g.sy = -p->Oy; g.ny = (int)g.Nyg-1-p->Oy; g.tz = (int)g.Nzg-1-p->Oz;
	g.has_w = g.wx>=0&&g.wx<(int)p->Nx; g.has_e = g.ex>=0&&g.ex<(int)p->Nx;
	g.has_s = g.sy>=0&&g.sy<(int)p->Ny; g.has_n = g.ny>=0&&g.ny<(int)p->Ny; g.has_t = g.tz>=0&&g.tz<(int)p->Nz;
	return g;
}

/* body force of the LUW step: global f + Coriolis (+ nudging + sponge), FX/kernel.cpp:1516-1614 */
static void luw_force(const luwo_params* p, const frame_t* g, uint64_t n, uint32_t bo, const float* u, int with_relaxation_zones,
	float rhon, float uxn, float uyn, float uzn, float fx, float fy, float fz, float ox, float oy, float oz, float* Fx, float* Fy, float* Fz) {
	const uint64_t N = cells(p);
	float fxn = fx, fyn = fy, fzn = fz;
	fxn += -2.0f*rhon*(oy*uzn-oz*uyn);
	fyn += -2.0f*rhon*(oz*uxn-ox*uzn);
	fzn += -2.0f*rhon*(ox*uyn-oy*uxn);
	if(with_relaxation_zones&&(p->features&LUWO_BUFFER_NUDGING)&&bo!=TYPE_E) {
		const xyz_t c = coords(p, n);
		const int xg = (int)c.x+p->Ox, yg = (int)c.y+p->Oy, zg = (int)c.z+p->Oz;
		const int Nb = (int)p->buffer_N;
		const int dw = xg, de = (int)(g->Nxg-1u)-xg, ds = yg, dn = (int)(g->Nyg-1u)-yg, dt = (int)(g->Nzg-1u)-zg;
		const int in_w = p->downstream_face!=1&&g->has_w&&dw>=0&&dw<=Nb;
		const int in_e = p->downstream_face!=2&&g->has_e&&de>=0&&de<=Nb;
		const int in_s = p->downstream_face!=3&&g->has_s&&ds>=0&&ds<=Nb;
		const int in_n = p->downstream_face!=4&&g->has_n&&dn>=0&&dn<=Nb;
		const int in_t = g->has_t&&dt>=0&&dt<=Nb;
		if(in_w||in_e||in_s||in_n||in_t) {
			uint32_t dmin = p->buffer_N+1u;
			uint64_t nref = n;
			if(in_w&&(uint32_t)dw<dmin) { dmin = (uint32_t)dw; nref = lin(p, (uint32_t)g->wx, c.y, c.z); }
			if(in_e&&(uint32_t)de<dmin) { dmin = (uint32_t)de; nref = lin(p, (uint32_t)g->ex, c.y, c.z); }
			if(in_s&&(uint32_t)ds<dmin) { dmin = (uint32_t)ds; nref = lin(p, c.x, (uint32_t)g->sy, c.z); }
			if(in_n&&(uint32_t)dn<dmin) { dmin = (uint32_t)dn; nref = lin(p, c.x, (uint32_t)g->ny, c.z); }
			if(in_t&&(uint32_t)dt<dmin) { dmin = (uint32_t)dt; nref = lin(p, c.x, c.y, (uint32_t)g->tz); }
			const float xi = 1.0f-(float)dmin/(float)p->buffer_N;
			float wb = sinf(1.5707963267948966f*xi);
			wb *= wb;
			const float ax = wb*p->buffer_inv_tau*(u[nref]-uxn);
			const float ay = wb*p->buffer_inv_tau*(u[N+nref]-uyn);
			const float az = p->buffer_nudge_vertical==1 ? wb*p->buffer_inv_tau*(u[2u*N+nref]-uzn) : 0.0f;
			fxn += rhon*ax; fyn += rhon*ay; fzn += rhon*az;
		}
	}
	if(with_relaxation_zones&&(p->features&LUWO_TOP_SPONGE)&&bo!=TYPE_E&&g->has_t) {
		const xyz_t c = coords(p, n);
		const int dt = (int)(g->Nzg-2u)-((int)c.z+p->Oz);
		const int Ns = (int)p->sponge_N;
		if(dt>=0&&dt<Ns) {
			const float xi = Ns>1 ? 1.0f-(float)dt/(float)(Ns-1) : 1.0f;
			float sigma = sinf(1.5707963267948966f*xi);
			sigma = p->sponge_inv_tau*sigma*sigma;
			const uint64_t nref = lin(p, c.x, c.y, (uint32_t)g->tz);
			fxn += rhon*sigma*(u[nref]-uxn);
			fyn += rhon*sigma*(u[N+nref]-uyn);
			fzn += rhon*sigma*(u[2u*N+nref]-uzn);
		}
	}
	*Fx = fxn; *Fy = fyn; *Fz = fzn;
}

/* ------------------------------------------------------------------ thermal D3Q7 (TEMPERATURE), FX/kernel.cpp:1306-1336 */
#define TYPE_T 0x04u
static void neighbors_temperature(const luwo_params* p, uint64_t n, uint64_t* j7) {
	const xyz_t c = coords(p, n);
	const uint64_t row = p->Nx, plane = (uint64_t)p->Nx*p->Ny;
	const uint64_t x0 = c.x, xp = (c.x+1u)%p->Nx, xm = (c.x+p->Nx-1u)%p->Nx;
	const uint64_t y0 = c.y*row, yp = ((c.y+1u)%p->Ny)*row, ym = ((c.y+p->Ny-1u)%p->Ny)*row;
	const uint64_t z0 = c.z*plane, zp = ((c.z+1u)%p->Nz)*plane, zm = ((c.z+p->Nz-1u)%p->Nz)*plane;
	j7[0] = n;
	j7[1] = xp+y0+z0; j7[2] = xm+y0+z0;
	j7[3] = x0+yp+z0; j7[4] = x0+ym+z0;
	j7[5] = x0+y0+zp; j7[6] = x0+y0+zm;
}
void luwo_calculate_g_eq(float T, float ux, float uy, float uz, float* geq) { /* D3Q7, lattice speed of sound 1/2, DDF-shifted */
	const float wsT4 = 0.5f*T, wsTm1 = 0.125f*(T-1.0f);
	geq[0] = fmaf(0.25f, T, -0.25f);
	geq[1] = fmaf(wsT4, ux, wsTm1); geq[2] = fmaf(wsT4, -ux, wsTm1);
	geq[3] = fmaf(wsT4, uy, wsTm1); geq[4] = fmaf(wsT4, -uy, wsTm1);
	geq[5] = fmaf(wsT4, uz, wsTm1); geq[6] = fmaf(wsT4, -uz, wsTm1);
}
static void load_g(const luwo_params* p, uint64_t n, float* g, const void* gi, const uint64_t* j7, uint64_t t) {
	const uint64_t N = cells(p);
	const int odd = (int)(t&1u);
	g[0] = ddf_load(p, gi, n);
	for(uint32_t i=1u; i<7u; i+=2u) {
		g[i   ] = ddf_load(p, gi, (uint64_t)(odd ? i    : i+1u)*N+n    );
		g[i+1u] = ddf_load(p, gi, (uint64_t)(odd ? i+1u : i   )*N+j7[i]);
	}
}
static void store_g(const luwo_params* p, uint64_t n, const float* g, void* gi, const uint64_t* j7, uint64_t t) {
	const uint64_t N = cells(p);
	const int odd = (int)(t&1u);
	ddf_store(p, gi, n, g[0]);
	for(uint32_t i=1u; i<7u; i+=2u) {
		ddf_store(p, gi, (uint64_t)(odd ? i+1u : i   )*N+j7[i], g[i   ]);
		ddf_store(p, gi, (uint64_t)(odd ? i    : i+1u)*N+n    , g[i+1u]);
	}
}
/* the TEMPERATURE block of stream_collide (FX/kernel.cpp:1639-1684): stream g in, T from g (or the preset of a TYPE_T cell), relax T towards the top
 * row inside the sponge (def_sponge_ref_mode == 0, the only mode LUW's build selects: clshim / FX/lbm.cpp:781), collide towards g_eq(T, u BEFORE the
 * force half-step), stream out, buoyancy onto the force. NOTE: the sponge reads T of the top-row cell of the column while that cell may store its
 * own T in the same NDRange unless it is TYPE_T; the reference has the same ordering freedom, so parity cases give the top row TYPE_T. */
static void thermal_step(const luwo_params* p, const luwo_thermal* th, const frame_t* g, uint64_t n, uint32_t fl, uint32_t bo, void* gi, float* T, uint64_t t,
	float uxn, float uyn, float uzn, float fx, float fy, float fz, float* fxn, float* fyn, float* fzn) {
	uint64_t j7[7];
	neighbors_temperature(p, n, j7);
	float gh[7];
	load_g(p, n, gh, gi, j7, t);
	float Tn;
	if(fl&TYPE_T) Tn = T[n];
	else {
		Tn = 0.0f;
		for(int i=0; i<7; i++) Tn += gh[i];
		Tn += 1.0f;
	}
	if((p->features&LUWO_TOP_SPONGE)&&!(fl&TYPE_T)&&bo!=TYPE_E&&g->has_t) {
		const xyz_t c = coords(p, n);
		const int dt = (int)(g->Nzg-2u)-((int)c.z+p->Oz);
		const int Ns = (int)p->sponge_N;
		if(dt>=0&&dt<Ns) {
			const float xi = Ns>1 ? 1.0f-(float)dt/(float)(Ns-1) : 1.0f;
			float sigma = sinf(1.5707963267948966f*xi);
			sigma = p->sponge_inv_tau*sigma*sigma;
			Tn = fmaf(sigma, T[lin(p, c.x, c.y, (uint32_t)g->tz)]-Tn, Tn);
		}
	}
	float geq[7];
	luwo_calculate_g_eq(Tn, uxn, uyn, uzn, geq);
	if(fl&TYPE_T) { for(int i=0; i<7; i++) gh[i] = geq[i]; }
	else {
		if(p->features&LUWO_UPDATE_FIELDS) T[n] = Tn;
		for(int i=0; i<7; i++) gh[i] = fmaf(1.0f-th->w_T, gh[i], th->w_T*geq[i]);
	}
	store_g(p, n, gh, gi, j7, t);
	*fxn -= fx*th->beta*(Tn-th->T_avg);
	*fyn -= fy*th->beta*(Tn-th->T_avg);
	*fzn -= fz*th->beta*(Tn-th->T_avg);
}

/* ------------------------------------------------------------------ kernel: initialize, FX/kernel.cpp:1370-1452 */
static void initialize_impl(const luwo_params* p, void* fi, const float* rho, float* u, uint8_t* flags, void* gi, const float* T) {
	const uint64_t N = cells(p);
	PAR_FOR
	for(int64_t nn=0; nn<(int64_t)N; nn++) {
		const uint64_t n = (uint64_t)nn;
		if(is_halo(p, coords(p, n))) continue;
		uint64_t j[Q];
		neighbors(p, n, j);
		if((flags[n]&TYPE_BO)==TYPE_S) { u[n] = 0.0f; u[N+n] = 0.0f; u[2u*N+n] = 0.0f; } /* MOVING_BOUNDARIES is off in LUW's build */
		float feq[Q];
		luwo_calculate_f_eq(rho[n], u[n], u[N+n], u[2u*N+n], feq);
		if(gi) { /* FX/kernel.cpp:1442-1450 */
			float geq[7];
			luwo_calculate_g_eq(T[n], u[n], u[N+n], u[2u*N+n], geq);
			uint64_t j7[7];
			neighbors_temperature(p, n, j7);
			store_g(p, n, geq, gi, j7, 1u);
		}
		store_f(p, n, feq, fi, j, 1u);
	}
}
void luwo_initialize(const luwo_params* p, void* fi, const float* rho, float* u, uint8_t* flags) { initialize_impl(p, fi, rho, u, flags, 0, 0); }
void luwo_initialize_thermal(const luwo_params* p, void* fi, const float* rho, float* u, uint8_t* flags, void* gi, const float* T) {
	initialize_impl(p, fi, rho, u, flags, gi, T);
}

/* ------------------------------------------------------------------ kernel: stream_collide, FX/kernel.cpp:1475-1780 */
static void stream_collide_impl(const luwo_params* p, const luwo_thermal* th, void* fi, float* rho, float* u, const uint8_t* flags, uint64_t t,
	float fx, float fy, float fz, float ox, float oy, float oz, void* gi, float* T) {
	const uint64_t N = cells(p);
	const frame_t g = frame(p);
	const int eq_on = (p->features&LUWO_EQUILIBRIUM_BOUNDARIES)!=0u;
	PAR_FOR
	for(int64_t nn=0; nn<(int64_t)N; nn++) {
		const uint64_t n = (uint64_t)nn;
		if(is_halo(p, coords(p, n))) continue;
		const uint32_t fl = flags[n], bo = fl&TYPE_BO, su = fl&TYPE_SU;
		if(bo==TYPE_S||su==TYPE_G) continue;
		uint64_t j[Q];
		neighbors(p, n, j);
		float f[Q];
		load_f(p, n, f, fi, j, t);
		const int is_e = eq_on&&bo==TYPE_E;
		float rhon, uxn, uyn, uzn;
		if(is_e) { rhon = rho[n]; uxn = u[n]; uyn = u[N+n]; uzn = u[2u*N+n]; }
		else rho_u(f, &rhon, &uxn, &uyn, &uzn);
		float fxn, fyn, fzn;
		luw_force(p, &g, n, bo, u, 1, rhon, uxn, uyn, uzn, fx, fy, fz, ox, oy, oz, &fxn, &fyn, &fzn);
		if(gi) thermal_step(p, th, &g, n, fl, bo, gi, T, t, uxn, uyn, uzn, fx, fy, fz, &fxn, &fyn, &fzn);
		float Fin[Q];
		if(p->features&LUWO_VOLUME_FORCE) {
			const float rho2 = 0.5f/rhon;
			uxn = clampf(fmaf(fxn, rho2, uxn), -LAT_C, LAT_C);
			uyn = clampf(fmaf(fyn, rho2, uyn), -LAT_C, LAT_C);
			uzn = clampf(fmaf(fzn, rho2, uzn), -LAT_C, LAT_C);
			forcing_terms(uxn, uyn, uzn, fxn, fyn, fzn, Fin);
		} else {
			uxn = clampf(uxn, -LAT_C, LAT_C); uyn = clampf(uyn, -LAT_C, LAT_C); uzn = clampf(uzn, -LAT_C, LAT_C);
			for(int i=0; i<Q; i++) Fin[i] = 0.0f;
		}
		if((p->features&LUWO_UPDATE_FIELDS)&&!is_e) { rho[n] = rhon; u[n] = uxn; u[N+n] = uyn; u[2u*N+n] = uzn; }
		float feq[Q];
		luwo_calculate_f_eq(rhon, uxn, uyn, uzn, feq);
		float w = p->w;
		if(p->features&LUWO_SUBGRID) { /* Smagorinsky-Lilly, FX/kernel.cpp:1723-1736 */
			const float tau0 = 1.0f/w;
			float Hxx=0.0f, Hyy=0.0f, Hzz=0.0f, Hxy=0.0f, Hxz=0.0f, Hyz=0.0f;
			for(int i=1; i<Q; i++) {
				const float fneq = f[i]-feq[i];
				Hxx += CX[i]*CX[i]*fneq;
				Hxy += CX[i]*CY[i]*fneq; Hyy += CY[i]*CY[i]*fneq;
				Hxz += CX[i]*CZ[i]*fneq; Hyz += CY[i]*CZ[i]*fneq; Hzz += CZ[i]*CZ[i]*fneq;
			}
			const float Qn = Hxx*Hxx+Hyy*Hyy+Hzz*Hzz+2.0f*(Hxy*Hxy+Hxz*Hxz+Hyz*Hyz);
			w = 2.0f/(tau0+sqrtf(tau0*tau0+0.76421222f*sqrtf(Qn)/rhon));
		}
		if(p->features&LUWO_VOLUME_FORCE) {
			const float c_tau = fmaf(w, -0.5f, 1.0f);
			for(int i=0; i<Q; i++) Fin[i] *= c_tau;
		}
		for(int i=0; i<Q; i++) f[i] = is_e ? feq[i] : fmaf(1.0f-w, f[i], fmaf(w, feq[i], Fin[i]));
		store_f(p, n, f, fi, j, t);
	}
}
void luwo_stream_collide(const luwo_params* p, void* fi, float* rho, float* u, const uint8_t* flags, uint64_t t,
	float fx, float fy, float fz, float ox, float oy, float oz) {
	stream_collide_impl(p, 0, fi, rho, u, flags, t, fx, fy, fz, ox, oy, oz, 0, 0);
}
void luwo_stream_collide_thermal(const luwo_params* p, const luwo_thermal* th, void* fi, float* rho, float* u, const uint8_t* flags, uint64_t t,
	float fx, float fy, float fz, float ox, float oy, float oz, void* gi, float* T) {
	stream_collide_impl(p, th, fi, rho, u, flags, t, fx, fy, fz, ox, oy, oz, gi, T);
}

/* ------------------------------------------------------------------ kernel: update_fields, FX/kernel.cpp:1938-2028 */
static void update_fields_impl(const luwo_params* p, const luwo_thermal* th, const void* fi, float* rho, float* u, const uint8_t* flags, uint64_t t,
	float fx, float fy, float fz, float ox, float oy, float oz, const void* gi, float* T) {
	const uint64_t N = cells(p);
	const frame_t g = frame(p);
	const int eq_on = (p->features&LUWO_EQUILIBRIUM_BOUNDARIES)!=0u;
	PAR_FOR
	for(int64_t nn=0; nn<(int64_t)N; nn++) {
		const uint64_t n = (uint64_t)nn;
		if(is_halo(p, coords(p, n))) continue;
		const uint32_t fl = flags[n], bo = fl&TYPE_BO, su = fl&TYPE_SU;
		if(bo==TYPE_S||su==TYPE_G) continue;
		uint64_t j[Q];
		neighbors(p, n, j);
		float f[Q];
		load_f(p, n, f, fi, j, t);
		float rhon, uxn, uyn, uzn;
		rho_u(f, &rhon, &uxn, &uyn, &uzn);
		float fxn, fyn, fzn;
		luw_force(p, &g, n, bo, u, 0, rhon, uxn, uyn, uzn, fx, fy, fz, ox, oy, oz, &fxn, &fyn, &fzn); /* no nudging/sponge in this kernel */
		if(gi) { /* FX/kernel.cpp:1981-2000: T from the streamed-in g (no sponge, no collision), buoyancy */
			uint64_t j7[7];
			neighbors_temperature(p, n, j7);
			float gh[7];
			load_g(p, n, gh, gi, j7, t);
			float Tn;
			if(fl&TYPE_T) Tn = T[n];
			else {
				Tn = 0.0f;
				for(int i=0; i<7; i++) Tn += gh[i];
				Tn += 1.0f;
				T[n] = Tn;
			}
			fxn -= fx*th->beta*(Tn-th->T_avg);
			fyn -= fy*th->beta*(Tn-th->T_avg);
			fzn -= fz*th->beta*(Tn-th->T_avg);
		}
		if(p->features&LUWO_VOLUME_FORCE) {
			const float rho2 = 0.5f/rhon;
			uxn = clampf(fmaf(fxn, rho2, uxn), -LAT_C, LAT_C);
			uyn = clampf(fmaf(fyn, rho2, uyn), -LAT_C, LAT_C);
			uzn = clampf(fmaf(fzn, rho2, uzn), -LAT_C, LAT_C);
		} else {
			uxn = clampf(uxn, -LAT_C, LAT_C); uyn = clampf(uyn, -LAT_C, LAT_C); uzn = clampf(uzn, -LAT_C, LAT_C);
		}
		if(!(eq_on&&bo==TYPE_E)) { rho[n] = rhon; u[n] = uxn; u[N+n] = uyn; u[2u*N+n] = uzn; }
	}
}
void luwo_update_fields(const luwo_params* p, const void* fi, float* rho, float* u, const uint8_t* flags, uint64_t t,
	float fx, float fy, float fz, float ox, float oy, float oz) {
	update_fields_impl(p, 0, fi, rho, u, flags, t, fx, fy, fz, ox, oy, oz, 0, 0);
}
void luwo_update_fields_thermal(const luwo_params* p, const luwo_thermal* th, const void* fi, float* rho, float* u, const uint8_t* flags, uint64_t t,
	float fx, float fy, float fz, float ox, float oy, float oz, const void* gi, float* T) {
	update_fields_impl(p, th, fi, rho, u, flags, t, fx, fy, fz, ox, oy, oz, gi, T);
}

/* ------------------------------------------------------------------ halo kernels, FX/kernel.cpp:2188-2297 */
static const uint8_t XFER[6][5] = { /* index_transfer(), D3Q19 */
	{1, 7,13, 9,15}, {2, 8,14,10,16}, {3, 7,14,11,17}, {4, 8,13,12,18}, {5, 9,16,11,18}, {6,10,15,12,17}
};
static inline uint64_t area(const luwo_params* p, uint32_t d) {
	return d==0u ? (uint64_t)p->Ny*p->Nz : d==1u ? (uint64_t)p->Nz*p->Nx : (uint64_t)p->Nx*p->Ny;
}
/* face cell a -> cell index on layer `layer` of axis `d` (index_extract_p/m, index_insert_p/m) */
static inline uint64_t face_cell(const luwo_params* p, uint32_t d, uint32_t a, uint32_t layer) {
	switch(d) {
		case 0u: return lin(p, layer, a%p->Ny, a/p->Ny);
		case 1u: return lin(p, a/p->Nz, layer, a%p->Nz);
		default: return lin(p, a%p->Nx, a/p->Nx, layer);
	}
}
static inline size_t ddf_size(const luwo_params* p) { return p->precision==LUWO_FP32 ? 4u : 2u; }
static void copy_ddf(const luwo_params* p, void* dst, uint64_t di, const void* src, uint64_t si) {
	if(ddf_size(p)==4u) ((uint32_t*)dst)[di] = ((const uint32_t*)src)[si];
	else ((uint16_t*)dst)[di] = ((const uint16_t*)src)[si];
}
void luwo_transfer_extract_fi(const luwo_params* p, uint32_t d, uint64_t t, void* buf_p, void* buf_m, const void* fi) {
	const uint64_t A = area(p, d), N = cells(p);
	const uint32_t L = d==0u ? p->Nx : d==1u ? p->Ny : p->Nz;
	const int odd = (int)(t&1u);
	PAR_FOR
	for(int64_t aa=0; aa<(int64_t)A; aa++) {
		const uint32_t a = (uint32_t)aa;
		for(int side=0; side<2; side++) {
			const uint64_t n = face_cell(p, d, a, side==0 ? L-2u : 1u);
			uint64_t j[Q];
			neighbors(p, n, j);
			for(uint32_t b=0u; b<5u; b++) {
				const uint32_t i = XFER[2u*d+(uint32_t)side][b];
				const uint64_t cell = (i&1u) ? j[i] : n;
				const uint32_t slot = odd ? ((i&1u) ? i+1u : i-1u) : i;
				copy_ddf(p, side==0 ? buf_p : buf_m, (uint64_t)b*A+a, fi, (uint64_t)slot*N+cell);
			}
		}
	}
}
void luwo_transfer_insert_fi(const luwo_params* p, uint32_t d, uint64_t t, const void* buf_p, const void* buf_m, void* fi) {
	const uint64_t A = area(p, d), N = cells(p);
	const uint32_t L = d==0u ? p->Nx : d==1u ? p->Ny : p->Nz;
	const int odd = (int)(t&1u);
	PAR_FOR
	for(int64_t aa=0; aa<(int64_t)A; aa++) {
		const uint32_t a = (uint32_t)aa;
		for(int side=0; side<2; side++) {
			const uint64_t n = face_cell(p, d, a, side==0 ? L-1u : 0u);
			uint64_t j[Q];
			neighbors(p, n, j);
			for(uint32_t b=0u; b<5u; b++) {
				const uint32_t i = XFER[2u*d+(uint32_t)side][b];
				const uint64_t cell = (i&1u) ? n : j[i-1u];
				const uint32_t slot = odd ? i : ((i&1u) ? i+1u : i-1u);
				copy_ddf(p, fi, (uint64_t)slot*N+cell, side==0 ? buf_p : buf_m, (uint64_t)b*A+a);
			}
		}
	}
}
void luwo_transfer_extract_rho_u_flags(const luwo_params* p, uint32_t d, char* buf_p, char* buf_m, const float* rho, const float* u, const uint8_t* flags) {
	const uint64_t A = area(p, d), N = cells(p);
	const uint32_t L = d==0u ? p->Nx : d==1u ? p->Ny : p->Nz;
	PAR_FOR
	for(int64_t aa=0; aa<(int64_t)A; aa++) {
		const uint32_t a = (uint32_t)aa;
		for(int side=0; side<2; side++) {
			const uint64_t n = face_cell(p, d, a, side==0 ? L-2u : 1u);
			char* buf = side==0 ? buf_p : buf_m;
			((float*)buf)[a] = rho[n]; ((float*)buf)[A+a] = u[n]; ((float*)buf)[2u*A+a] = u[N+n]; ((float*)buf)[3u*A+a] = u[2u*N+n];
			((uint8_t*)buf)[16u*A+a] = flags[n];
		}
	}
}
void luwo_transfer_insert_rho_u_flags(const luwo_params* p, uint32_t d, const char* buf_p, const char* buf_m, float* rho, float* u, uint8_t* flags) {
	const uint64_t A = area(p, d), N = cells(p);
	const uint32_t L = d==0u ? p->Nx : d==1u ? p->Ny : p->Nz;
	PAR_FOR
	for(int64_t aa=0; aa<(int64_t)A; aa++) {
		const uint32_t a = (uint32_t)aa;
		for(int side=0; side<2; side++) {
			const uint64_t n = face_cell(p, d, a, side==0 ? L-1u : 0u);
			const char* buf = side==0 ? buf_p : buf_m;
			rho[n] = ((const float*)buf)[a]; u[n] = ((const float*)buf)[A+a]; u[N+n] = ((const float*)buf)[2u*A+a]; u[2u*N+n] = ((const float*)buf)[3u*A+a];
			flags[n] = ((const uint8_t*)buf)[16u*A+a];
		}
	}
}

/* thermal halos, FX/kernel.cpp:2337-2377: one g DDF per face cell and side (i = 2*direction+1 towards +, 2*direction+2 towards -), and the T field */
void luwo_transfer_extract_gi(const luwo_params* p, uint32_t d, uint64_t t, void* buf_p, void* buf_m, const void* gi) {
	const uint64_t A = area(p, d), N = cells(p);
	const uint32_t L = d==0u ? p->Nx : d==1u ? p->Ny : p->Nz;
	const int odd = (int)(t&1u);
	PAR_FOR
	for(int64_t aa=0; aa<(int64_t)A; aa++) {
		const uint32_t a = (uint32_t)aa;
		for(int side=0; side<2; side++) {
			const uint64_t n = face_cell(p, d, a, side==0 ? L-2u : 1u);
			uint64_t j7[7];
			neighbors_temperature(p, n, j7);
			const uint32_t i = 2u*d+(uint32_t)side+1u;
			const uint64_t cell = (i&1u) ? j7[i] : n;
			const uint32_t slot = odd ? ((i&1u) ? i+1u : i-1u) : i;
			copy_ddf(p, side==0 ? buf_p : buf_m, a, gi, (uint64_t)slot*N+cell);
		}
	}
}
void luwo_transfer_insert_gi(const luwo_params* p, uint32_t d, uint64_t t, const void* buf_p, const void* buf_m, void* gi) {
	const uint64_t A = area(p, d), N = cells(p);
	const uint32_t L = d==0u ? p->Nx : d==1u ? p->Ny : p->Nz;
	const int odd = (int)(t&1u);
	PAR_FOR
	for(int64_t aa=0; aa<(int64_t)A; aa++) {
		const uint32_t a = (uint32_t)aa;
		for(int side=0; side<2; side++) {
			const uint64_t n = face_cell(p, d, a, side==0 ? L-1u : 0u);
			uint64_t j7[7];
			neighbors_temperature(p, n, j7);
			const uint32_t i = 2u*d+(uint32_t)side+1u;
			const uint64_t cell = (i&1u) ? n : j7[i-1u];
			const uint32_t slot = odd ? i : ((i&1u) ? i+1u : i-1u);
			copy_ddf(p, gi, (uint64_t)slot*N+cell, side==0 ? buf_p : buf_m, a);
		}
	}
}
void luwo_transfer_extract_T(const luwo_params* p, uint32_t d, float* buf_p, float* buf_m, const float* T) {
	const uint64_t A = area(p, d);
	const uint32_t L = d==0u ? p->Nx : d==1u ? p->Ny : p->Nz;
	PAR_FOR
	for(int64_t aa=0; aa<(int64_t)A; aa++) {
		const uint32_t a = (uint32_t)aa;
		buf_p[a] = T[face_cell(p, d, a, L-2u)];
		buf_m[a] = T[face_cell(p, d, a, 1u)];
	}
}
void luwo_transfer_insert_T(const luwo_params* p, uint32_t d, const float* buf_p, const float* buf_m, float* T) {
	const uint64_t A = area(p, d);
	const uint32_t L = d==0u ? p->Nx : d==1u ? p->Ny : p->Nz;
	PAR_FOR
	for(int64_t aa=0; aa<(int64_t)A; aa++) {
		const uint32_t a = (uint32_t)aa;
		T[face_cell(p, d, a, L-1u)] = buf_p[a];
		T[face_cell(p, d, a, 0u)] = buf_m[a];
	}
}

/* ------------------------------------------------------------------ kernel: vk_inlet_apply, FX/kernel.cpp:2495-2571 */
void luwo_vk_inlet_apply(uint64_t Ncells, uint32_t use_interp, float t0, float t1, float alpha, uint64_t P, uint64_t M, uint64_t V,
	const uint64_t* point_cell, const uint8_t* point_face, const float* pd, const float* md, float* u) {
	PAR_FOR
	for(int64_t ii=0; ii<(int64_t)P; ii++) {
		const uint64_t i = (uint64_t)ii, n = point_cell[i];
		const uint64_t fid = point_face[i]&0x07u;
		const float px = pd[i], py = pd[P+i], pz = pd[2u*P+i];
		const float ubx = pd[3u*P+i], uby = pd[4u*P+i], ubz = pd[5u*P+i], sigma = pd[6u*P+i];
		if(fid>=5u||!(sigma>0.0f)) { u[n] = ubx; u[Ncells+n] = uby; u[2u*Ncells+n] = ubz; continue; }
		float qx = 0.0f, qy = 0.0f, qz = 0.0f;
		for(uint64_t m=0u; m<M; m++) {
			const uint64_t k = fid*M+m;
			const float kx = md[k], ky = md[V+k], kz = md[2u*V+k], om = md[3u*V+k];
			const float Ax = md[4u*V+k], Ay = md[5u*V+k], Az = md[6u*V+k];
			const float phx = md[7u*V+k], phy = md[8u*V+k], phz = md[9u*V+k];
			const float ph0 = fmaf(kx, px, fmaf(ky, py, fmaf(kz, pz, om*t0)));
			float vx = Ax*cosf(ph0+phx), vy = Ay*cosf(ph0+phy), vz = Az*cosf(ph0+phz);
			if(use_interp!=0u) {
				const float ph1 = fmaf(kx, px, fmaf(ky, py, fmaf(kz, pz, om*t1)));
				const float vx1 = Ax*cosf(ph1+phx), vy1 = Ay*cosf(ph1+phy), vz1 = Az*cosf(ph1+phz);
				vx = fmaf(alpha, vx1-vx, vx); vy = fmaf(alpha, vy1-vy, vy); vz = fmaf(alpha, vz1-vz, vz);
			}
			qx += vx; qy += vy; qz += vz;
		}
		u[n] = fmaf(sigma, qx, ubx); u[Ncells+n] = fmaf(sigma, qy, uby); u[2u*Ncells+n] = fmaf(sigma, qz, ubz);
	}
}

/* FX/setup.cpp:4441-4488: ++avg_count; inv_n = 1/avg_count; per cell Welford update of the three velocity components and the running mean of rho */
void luwo_stats_accumulate(uint64_t N, uint64_t count, const float* rho, const float* u, float* u_avg, float* rho_avg, float* m2_u, float* m2_v, float* m2_w) {
	const float inv_n = 1.0f/(float)count;
	const float* ux = u; const float* uy = u+N; const float* uz = u+2u*N;
	int64_t n;
#pragma omp parallel for schedule(static) num_threads(luwo_get_threads())
	for(n=0; n<(int64_t)N; n++) {
		const uint64_t i3 = 3ull*(uint64_t)n;
		float mean_u = u_avg[i3], mean_v = u_avg[i3+1u], mean_w = u_avg[i3+2u];
		const float delta_u = ux[n]-mean_u; mean_u += delta_u*inv_n; m2_u[n] += delta_u*(ux[n]-mean_u); u_avg[i3] = mean_u;
		const float delta_v = uy[n]-mean_v; mean_v += delta_v*inv_n; m2_v[n] += delta_v*(uy[n]-mean_v); u_avg[i3+1u] = mean_v;
		const float delta_w = uz[n]-mean_w; mean_w += delta_w*inv_n; m2_w[n] += delta_w*(uz[n]-mean_w); u_avg[i3+2u] = mean_w;
		rho_avg[n] += (rho[n]-rho_avg[n])*inv_n;
	}
}

/* FX/kernel.cpp:2381-2471. Vector helpers written out in the reference's operation order (cross: a.y*b.z-a.z*b.y, ...; dot: left to right). */
static int clampi(int x, int lo, int hi) { return x<lo ? lo : (x>hi ? hi : x); }
void luwo_voxelize_mesh(const luwo_params* p, uint32_t direction, const float* u, uint8_t* flags, uint8_t flag, const float* p0, const float* p1, const float* p2, const float* bbu) {
	const uint32_t Nx = p->Nx, Ny = p->Ny, Nz = p->Nz;
	const uint64_t N = (uint64_t)Nx*Ny*Nz;
	const int Ox = p->Ox, Oy = p->Oy, Oz = p->Oz;
	const uint32_t A = direction==0u ? Ny*Nz : direction==1u ? Nz*Nx : Nx*Ny;
	uint32_t triangle_number; memcpy(&triangle_number, bbu, 4);
	const float x0 = bbu[1], y0 = bbu[2], z0 = bbu[3], x1 = bbu[4], y1 = bbu[5], z1 = bbu[6];
	const float offx = 0.5f*(float)((int)Nx+2*Ox)-0.5f, offy = 0.5f*(float)((int)Ny+2*Oy)-0.5f, offz = 0.5f*(float)((int)Nz+2*Oz)-0.5f;
	const float dx = (float)(direction==0u), dy = (float)(direction==1u), dz = (float)(direction==2u);
	int64_t ai;
#pragma omp parallel for schedule(dynamic, 64) num_threads(luwo_get_threads())
	for(ai=0; ai<(int64_t)A; ai++) {
		const uint32_t a = (uint32_t)ai;
		uint32_t X, Y, Z;
		if(direction==0u) { X = (uint32_t)clampi((int)x0-Ox, 0, (int)Nx-1); Y = a%Ny; Z = a/Ny; }
		else if(direction==1u) { X = a/Nz; Y = (uint32_t)clampi((int)y0-Oy, 0, (int)Ny-1); Z = a%Nz; }
		else { X = a%Nx; Y = a/Nx; Z = (uint32_t)clampi((int)z0-Oz, 0, (int)Nz-1); }
		const float rox = ((float)X+0.5f-0.5f*(float)Nx)+offx, roy = ((float)Y+0.5f-0.5f*(float)Ny)+offy, roz = ((float)Z+0.5f-0.5f*(float)Nz)+offz;
		const int out_of_box = direction==0u ? (roy<y0||roz<z0||roy>=y1||roz>=z1) : direction==1u ? (rox<x0||roz<z0||rox>=x1||roz>=z1) : (rox<x0||roy<y0||rox>=x1||roy>=y1);
		if(out_of_box) continue;
		uint32_t intersections = 0u, intersections_check = 0u;
		uint16_t distances[64];
		for(uint32_t i=0u; i<triangle_number; i++) {
			const float ax = p0[3u*i], ay = p0[3u*i+1u], az = p0[3u*i+2u];
			const float ux_ = p1[3u*i]-ax, uy_ = p1[3u*i+1u]-ay, uz_ = p1[3u*i+2u]-az;
			const float vx = p2[3u*i]-ax, vy = p2[3u*i+1u]-ay, vz = p2[3u*i+2u]-az;
			const float wx = rox-ax, wy = roy-ay, wz = roz-az;
			const float hx = dy*vz-dz*vy, hy = dz*vx-dx*vz, hz = dx*vy-dy*vx; /* cross(r_direction, v) */
			const float qx = wy*uz_-wz*uy_, qy = wz*ux_-wx*uz_, qz = wx*uy_-wy*ux_; /* cross(w, u) */
			const float g = ux_*hx+uy_*hy+uz_*hz, f = 1.0f/g, s = f*(wx*hx+wy*hy+wz*hz), t = f*(dx*qx+dy*qy+dz*qz), d = f*(vx*qx+vy*qy+vz*qz);
			if(g!=0.0f&&s>=0.0f&&s<1.0f&&t>=0.0f&&s+t<1.0f) {
				if(d>0.0f) { if(intersections<64u&&d<65536.0f) distances[intersections] = (uint16_t)d; intersections++; }
				else intersections_check++;
			}
		}
		const uint32_t nsort = intersections<64u ? intersections : 64u;
		for(uint32_t i=1u; i<nsort; i++) { const uint16_t t = distances[i]; int j = (int)i-1; while(j>=0&&distances[j]>t) { distances[j+1] = distances[j]; j--; } distances[j+1] = t; }
		int inside = (intersections%2u)&&(intersections_check%2u);
		uint32_t intersection = intersections%2u!=intersections_check%2u;
		const uint32_t h0 = direction==0u ? X : direction==1u ? Y : Z;
		const uint32_t hmax = direction==0u ? (uint32_t)clampi((int)x1-Ox, 0, (int)Nx) : direction==1u ? (uint32_t)clampi((int)y1-Oy, 0, (int)Ny) : (uint32_t)clampi((int)z1-Oz, 0, (int)Nz);
		const uint32_t last = intersections-1u<63u ? intersections-1u : 63u; /* min(intersections-1u, 63u) with unsigned wrap for 0 */
		const uint32_t hmesh = h0+(uint32_t)(intersections>0u ? distances[last] : 0u); /* intersections == 0: the reference reads an uninitialised slot, but `inside` is false then and stays false */
		for(uint32_t h=h0; h<hmax; h++) {
			while(intersection<intersections&&h>h0+(uint32_t)distances[intersection<63u ? intersection : 63u]) { inside = !inside; intersection++; }
			inside = inside&&(intersection<intersections&&h<hmesh);
			const uint64_t n = (uint64_t)(direction==0u ? h : X)+((uint64_t)(direction==1u ? h : Y)+(uint64_t)(direction==2u ? h : Z)*Ny)*Nx;
			uint8_t fl = flags[n];
			if(inside) fl = (uint8_t)((fl&~TYPE_BO)|flag);
			else if((fl&TYPE_BO)==TYPE_S) {
				if(u[n]==0.0f&&u[N+n]==0.0f&&u[2u*N+n]==0.0f) fl = (fl&TYPE_BO)==TYPE_BO ? (uint8_t)(fl&~TYPE_BO) : (uint8_t)(fl&~flag); /* u_set == 0 for resting geometry; TYPE_MS == TYPE_BO == 0x03 */
			}
			flags[n] = fl;
		}
	}
}
