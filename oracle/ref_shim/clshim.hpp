// TEST INFRASTRUCTURE (oracle/): a minimal emulation of the OpenCL-C environment the reference's LBM kernels need,
// so that the reference kernel text (extracted by make_ref_kernel.py into oracle/_ref/kernel_cl.inc) can be compiled
// by g++ and executed on host threads. Written from scratch; contains no reference code.
//
// * Every `def_*` constant that the reference bakes into the JIT source per case (FX/lbm.cpp:612-783) is mapped to a
//   runtime global here, so one shared object serves every grid / decomposition.
// * Feature switches (UPDATE_FIELDS, SUBGRID, BUFFER_NUDGING, ...) and the DDF precision (FP16S / FP16C / FP32)
//   stay compile-time, exactly as in the reference; the Makefile builds one .so per combination.
// * Floating point: compile with -ffp-contract=off so that only the reference's explicit fma() calls are fused
//   ("as written" semantics; OpenCL's -cl-mad-enable *permits* but does not require contraction).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

// OpenCL C math built-ins are generic over float: make the float overloads visible at global scope. Without this, fma(float,float,float)
// would bind to ::fma(double,double,double) from <math.h> and round twice.
using std::fma; using std::sqrt; using std::sin; using std::cos; using std::tan; using std::asin; using std::acos; using std::atan; using std::atan2;
using std::fabs; using std::fmin; using std::fmax; using std::fmod; using std::floor; using std::ceil; using std::round; using std::exp; using std::log;
using std::pow; using std::isnan; using std::isinf; using std::isfinite; using std::copysign; using std::cbrt; using std::exp2; using std::log2;

typedef unsigned char uchar;
typedef unsigned short ushort;
typedef unsigned int uint;
typedef unsigned long ulong; // 64-bit on LP64, like OpenCL ulong
static_assert(sizeof(ulong)==8, "LP64 required");

// ---------------------------------------------------------------- address-space / function qualifiers
#define kernel
#define global
#define local
#define constant const
#define __kernel
#define __global
#define __local
#define __constant const

// ---------------------------------------------------------------- work-item id
static thread_local ulong cl_gid = 0ul;
#define get_global_id(d) (cl_gid)

// ---------------------------------------------------------------- vector types (only what the non-graphics code touches)
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct int3 { int x, y, z; };
struct uint3 { uint x, y, z; };
struct uchar4 { uchar x, y, z, w; };
static inline float2 cl_make_float2(float x, float y) { return float2{x, y}; }
static inline float3 cl_make_float3(float x, float y, float z) { return float3{x, y, z}; }
static inline float3 cl_make_float3(float s) { return float3{s, s, s}; }
static inline int3 cl_make_int3(int x, int y, int z) { return int3{x, y, z}; }
static inline uint3 cl_make_uint3(uint x, uint y, uint z) { return uint3{x, y, z}; }
static inline uchar4 cl_make_uchar4(uchar x, uchar y, uchar z, uchar w) { return uchar4{x, y, z, w}; }
static inline float3 operator+(const float3 a, const float3 b) { return float3{a.x+b.x, a.y+b.y, a.z+b.z}; }
static inline float3 operator-(const float3 a, const float3 b) { return float3{a.x-b.x, a.y-b.y, a.z-b.z}; }
static inline float3 operator-(const float3 a) { return float3{-a.x, -a.y, -a.z}; }
static inline float3 operator*(const float s, const float3 a) { return float3{s*a.x, s*a.y, s*a.z}; }
static inline float3 operator*(const float3 a, const float s) { return float3{a.x*s, a.y*s, a.z*s}; }
static inline float3 operator*(const float3 a, const float3 b) { return float3{a.x*b.x, a.y*b.y, a.z*b.z}; }
static inline float3 operator/(const float3 a, const float s) { return float3{a.x/s, a.y/s, a.z/s}; }
static inline float3& operator+=(float3& a, const float3 b) { a = a+b; return a; }
static inline float3& operator-=(float3& a, const float3 b) { a = a-b; return a; }
static inline float3& operator*=(float3& a, const float s) { a = a*s; return a; }
static inline float dot(const float3 a, const float3 b) { return a.x*b.x+a.y*b.y+a.z*b.z; }
static inline float3 cross(const float3 a, const float3 b) { return float3{a.y*b.z-a.z*b.y, a.z*b.x-a.x*b.z, a.x*b.y-a.y*b.x}; }
static inline float length(const float3 a) { return sqrtf(dot(a, a)); }
static inline float3 normalize(const float3 a) { return a/length(a); }

// ---------------------------------------------------------------- scalar built-ins
static inline uint as_uint(const float x) { uint r; memcpy(&r, &x, 4); return r; }
static inline int as_int(const float x) { int r; memcpy(&r, &x, 4); return r; }
static inline float as_float(const uint x) { float r; memcpy(&r, &x, 4); return r; }
static inline float as_float(const int x) { float r; memcpy(&r, &x, 4); return r; }
static inline float clamp(const float x, const float lo, const float hi) { return fminf(fmaxf(x, lo), hi); }
static inline int clamp(const int x, const int lo, const int hi) { return x<lo ? lo : (x>hi ? hi : x); }
static inline float sign(const float x) { return x>0.0f ? 1.0f : (x<0.0f ? -1.0f : 0.0f); }
static inline float rsqrt(const float x) { return 1.0f/sqrtf(x); }
static inline float native_rsqrt(const float x) { return 1.0f/sqrtf(x); }
static inline float native_sqrt(const float x) { return sqrtf(x); }
template<typename T> static inline T min(const T a, const T b) { return a<b ? a : b; }
template<typename T> static inline T max(const T a, const T b) { return a>b ? a : b; }

// IEEE-754 binary16 <-> binary32 (round-to-nearest-even, subnormals, inf/nan) -- what vload_half / vstore_half_rte do in OpenCL C
static inline float cl_half_bits_to_float(const ushort h) {
	const uint s = (uint)(h&0x8000u)<<16;
	uint e = (h>>10)&0x1Fu, m = h&0x3FFu;
	if(e==0u) {
		if(m==0u) return as_float(s);
		int k = 0; // normalise subnormal
		while(!(m&0x400u)) { m <<= 1; k++; }
		m &= 0x3FFu;
		return as_float(s|((uint)(113-k)<<23)|(m<<13));
	}
	if(e==31u) return as_float(s|0x7F800000u|(m<<13));
	return as_float(s|((e+112u)<<23)|(m<<13));
}
static inline ushort cl_float_to_half_bits_rte(const float f) {
	const uint x = as_uint(f);
	const uint s = (x>>16)&0x8000u;
	const uint a = x&0x7FFFFFFFu;
	if(a>=0x7F800000u) return (ushort)(s|0x7C00u|(a>0x7F800000u ? 0x200u|((a>>13)&0x3FFu) : 0u)); // inf / nan
	if(a>=0x477FF000u) return (ushort)(s|0x7C00u); // >= 65520 rounds to inf
	if(a<0x33000001u) return (ushort)s; // <= 2^-25 rounds to zero (ties-to-even at exactly 2^-25)
	int e = (int)(a>>23)-127;
	uint m = (a&0x007FFFFFu)|0x00800000u; // 24-bit significand
	int shift; // number of low bits to drop
	uint he;
	if(e<-14) { shift = 13+(-14-e); he = 0u; } else { shift = 13; he = (uint)(e+15); }
	const uint keep = m>>shift;
	const uint rem = m&((1u<<shift)-1u);
	const uint half = 1u<<(shift-1);
	uint r = (he==0u ? keep : (keep&0x3FFu))|(he<<10);
	if(rem>half||(rem==half&&(keep&1u))) r++; // carries propagate into the exponent correctly
	return (ushort)(s|r);
}
static inline float vload_half(const ulong o, const ushort* p) { return cl_half_bits_to_float(p[o]); }
static inline void vstore_half_rte(const float x, const ulong o, ushort* p) { p[o] = cl_float_to_half_bits_rte(x); }

// ---------------------------------------------------------------- per-case constants -> runtime globals (FX/lbm.cpp:612-783)
static uint g_Nx=1u, g_Ny=1u, g_Nz=1u; static ulong g_N=1ul;
static uint g_Dx=1u, g_Dy=1u, g_Dz=1u; static int g_Ox=0, g_Oy=0, g_Oz=0;
static uint g_Nx_global=1u, g_Ny_global=1u, g_Nz_global=1u;
static int g_west_local_x=0, g_east_local_x=0, g_south_local_y=0, g_north_local_y=0, g_top_local_z=0;
static int g_has_west=0, g_has_east=0, g_has_south=0, g_has_north=0, g_has_top=0;
static float g_w=1.0f, g_w_T=1.0f, g_beta=0.0f, g_T_avg=1.0f;
static int g_downstream_face=0, g_buffer_nudge_vertical=0;
static uint g_buffer_N=1u, g_sponge_N=1u;
static float g_buffer_inv_tau=0.0f, g_sponge_inv_tau=0.0f;

#define def_Nx g_Nx
#define def_Ny g_Ny
#define def_Nz g_Nz
#define def_N g_N
#define uxx uint
#define def_Dx g_Dx
#define def_Dy g_Dy
#define def_Dz g_Dz
#define def_Ox g_Ox
#define def_Oy g_Oy
#define def_Oz g_Oz
#define def_Nx_global g_Nx_global
#define def_Ny_global g_Ny_global
#define def_Nz_global g_Nz_global
#define def_west_local_x g_west_local_x
#define def_east_local_x g_east_local_x
#define def_south_local_y g_south_local_y
#define def_north_local_y g_north_local_y
#define def_top_local_z g_top_local_z
#define def_has_west_face g_has_west
#define def_has_east_face g_has_east
#define def_has_south_face g_has_south
#define def_has_north_face g_has_north
#define def_has_top_face g_has_top
#define def_Ax (g_Ny*g_Nz)
#define def_Ay (g_Nz*g_Nx)
#define def_Az (g_Nx*g_Ny)
#define D3Q19
#define def_velocity_set 19u
#define def_dimensions 3u
#define def_transfers 5u
#define def_c 0.57735027f
#define def_w g_w
#define def_w0 (1.0f/3.0f)
#define def_ws (1.0f/18.0f)
#define def_we (1.0f/36.0f)
#define SRT
#define TYPE_S 0x01
#define TYPE_E 0x02
#define TYPE_T 0x04
#define TYPE_F 0x08
#define TYPE_I 0x10
#define TYPE_G 0x20
#define TYPE_X 0x40
#define TYPE_Y 0x80
#define TYPE_MS 0x03
#define TYPE_BO 0x03
#define TYPE_IF 0x18
#define TYPE_IG 0x30
#define TYPE_GI 0x38
#define TYPE_SU 0x38
#define def_w_T g_w_T
#define def_beta g_beta
#define def_T_avg g_T_avg
#define def_downstream_face g_downstream_face
#define def_buffer_N g_buffer_N
#define def_buffer_inv_tau g_buffer_inv_tau
#define def_buffer_nudge_vertical g_buffer_nudge_vertical
#define def_sponge_N g_sponge_N
#define def_sponge_inv_tau g_sponge_inv_tau
#define def_sponge_ref_mode 0

// DDF storage macros exactly as FX/lbm.cpp:706-721 defines them for the JIT
#if defined(FP16S)
#define fpxx ushort
#define fpxx_copy ushort
#define load(p,o) vload_half(o,p)*3.0517578E-5f
#define store(p,o,x) vstore_half_rte((x)*32768.0f,o,p)
#elif defined(FP16C)
#define fpxx ushort
#define fpxx_copy ushort
#define load(p,o) half_to_float_custom(p[o])
#define store(p,o,x) p[o]=float_to_half_custom(x)
#else
#define fpxx float
#define fpxx_copy float
#define load(p,o) p[o]
#define store(p,o,x) p[o]=x
#endif
