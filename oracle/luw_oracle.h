/* TEST INFRASTRUCTURE -- CPU oracle for the LatticeUrbanWind LBM time step.
 *
 * This is a plain-C restatement of the algorithm in the reference's OpenCL-C device code
 * (FX = /root/reference/core/cfd_core/FluidX3D/src): FX/kernel.cpp:833-1113 (indexing, codecs, f_eq, rho/u, Guo forcing),
 * :1338-1351 (Esoteric-Pull load/store), :1370-1452 (initialize), :1475-1780 (stream_collide), :1938-2028 (update_fields),
 * :2188-2310 (halo extract/insert), :2495-2571 (vk_inlet_apply); thermal D3Q7: :1306-1336, :1639-1684, :1981-2000, :2337-2377.
 *
 * It is the CHECKER, never the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it. The shipped CUDA path never links or calls anything in oracle/.
 *
 * PARITY PIN: the reference ships no golden vectors for this path. The oracle is instead pinned against the
 * reference's OWN kernel text, compiled for host threads through oracle/ref_shim (oracle/_ref/libluwref_*.so), bit for bit:
 * see tests/test_oracle_vs_reference.py and the committed fixtures in tests/golden/ produced by that build.
 *
 * Floating point: "as written" semantics. Compile with -ffp-contract=off; fused operations appear only where the
 * reference writes fma(). Division and sqrt are IEEE correctly rounded.
 */
#ifndef LUW_ORACLE_H
#define LUW_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { LUWO_FP32 = 0, LUWO_FP16S = 1, LUWO_FP16C = 2 };
enum { /* compile-time switches of the reference (FX/defines.hpp:17-29) as run-time bits */
	LUWO_UPDATE_FIELDS = 1u, LUWO_VOLUME_FORCE = 2u, LUWO_EQUILIBRIUM_BOUNDARIES = 4u, LUWO_SUBGRID = 8u,
	LUWO_BUFFER_NUDGING = 16u, LUWO_TOP_SPONGE = 32u
};
/* the TEMPERATURE extension is selected per call: the *_thermal entry points take the D3Q7 DDFs `gi` and the field `T` (TYPE_T = flag bit 0x04) */

typedef struct luwo_params { /* per-domain constants, FX/lbm.cpp:612-783 */
	uint32_t Nx, Ny, Nz; /* local lattice incl. halo layers */
	uint32_t Dx, Dy, Dz; /* domains per axis */
	int32_t Ox, Oy, Oz; /* global coordinate of local cell 0 */
	uint32_t precision; /* LUWO_FP32 / FP16S / FP16C */
	uint32_t features; /* LUWO_* bits */
	float w; /* def_w = 1/tau */
	int32_t downstream_face; /* 0 none, 1 west, 2 east, 3 south, 4 north */
	uint32_t buffer_N; float buffer_inv_tau; int32_t buffer_nudge_vertical;
	uint32_t sponge_N; float sponge_inv_tau;
} luwo_params;

typedef struct luwo_thermal { /* FX/lbm.cpp:750-752 */
	float w_T; /* def_w_T = 1/(2 alpha + 1/2) */
	float beta; /* def_beta: thermal expansion coefficient (buoyancy = -f * beta * (T - T_avg); LUW passes f = 0) */
	float T_avg; /* def_T_avg */
} luwo_thermal;

/* codecs (FX/kernel.cpp:864-875, FX/lbm.cpp:706-721) */
float luwo_half_to_float(uint16_t h); /* IEEE binary16 -> binary32 */
uint16_t luwo_float_to_half_rte(float f); /* IEEE binary32 -> binary16, round-to-nearest-even */
float luwo_fp16c_to_float(uint16_t x);
uint16_t luwo_float_to_fp16c(float x);

void luwo_calculate_f_eq(float rho, float ux, float uy, float uz, float* feq);

/* kernels: one call == one NDRange of the reference */
void luwo_initialize(const luwo_params* p, void* fi, const float* rho, float* u, uint8_t* flags);
void luwo_stream_collide(const luwo_params* p, void* fi, float* rho, float* u, const uint8_t* flags, uint64_t t,
	float fx, float fy, float fz, float omega_x, float omega_y, float omega_z);
void luwo_update_fields(const luwo_params* p, const void* fi, float* rho, float* u, const uint8_t* flags, uint64_t t,
	float fx, float fy, float fz, float omega_x, float omega_y, float omega_z);
void luwo_transfer_extract_fi(const luwo_params* p, uint32_t direction, uint64_t t, void* buf_p, void* buf_m, const void* fi);
void luwo_transfer_insert_fi(const luwo_params* p, uint32_t direction, uint64_t t, const void* buf_p, const void* buf_m, void* fi);
void luwo_transfer_extract_rho_u_flags(const luwo_params* p, uint32_t direction, char* buf_p, char* buf_m, const float* rho, const float* u, const uint8_t* flags);
void luwo_transfer_insert_rho_u_flags(const luwo_params* p, uint32_t direction, const char* buf_p, const char* buf_m, float* rho, float* u, uint8_t* flags);
void luwo_vk_inlet_apply(uint64_t N_cells, uint32_t use_interp, float t0, float t1, float alpha, uint64_t point_count, uint64_t mode_count, uint64_t mode_stride,
	const uint64_t* point_cell, const uint8_t* point_face, const float* point_data, const float* mode_data, float* u);

/* thermal D3Q7 extension (TEMPERATURE): FX/kernel.cpp:1306-1336 (g_eq, Esoteric-Pull load_g / store_g), :1442-1450 (initialize), :1639-1684 (the block inside
 * stream_collide: T, sponge on T, SRT collision of g, buoyancy), :1981-2000 (update_fields), :2337-2377 (halos of gi and T). gi: 7 x N DDFs in the same
 * storage type as fi; T: N floats. The momentum part of these entry points is the code of the plain ones. */
void luwo_calculate_g_eq(float T, float ux, float uy, float uz, float* geq);
void luwo_initialize_thermal(const luwo_params* p, void* fi, const float* rho, float* u, uint8_t* flags, void* gi, const float* T);
void luwo_stream_collide_thermal(const luwo_params* p, const luwo_thermal* th, void* fi, float* rho, float* u, const uint8_t* flags, uint64_t t,
	float fx, float fy, float fz, float omega_x, float omega_y, float omega_z, void* gi, float* T);
void luwo_update_fields_thermal(const luwo_params* p, const luwo_thermal* th, const void* fi, float* rho, float* u, const uint8_t* flags, uint64_t t,
	float fx, float fy, float fz, float omega_x, float omega_y, float omega_z, const void* gi, float* T);
void luwo_transfer_extract_gi(const luwo_params* p, uint32_t direction, uint64_t t, void* buf_p, void* buf_m, const void* gi);
void luwo_transfer_insert_gi(const luwo_params* p, uint32_t direction, uint64_t t, const void* buf_p, const void* buf_m, void* gi);
void luwo_transfer_extract_T(const luwo_params* p, uint32_t direction, float* buf_p, float* buf_m, const float* T);
void luwo_transfer_insert_T(const luwo_params* p, uint32_t direction, const float* buf_p, const float* buf_m, float* T);

/* kernel voxelize_mesh, FX/kernel.cpp:2381-2471, for resting geometry (bbu[10..15] == 0, the only way LUW calls it: FX/setup.cpp passes no velocities):
 * one ray per column of the face normal to `direction`, Moeller-Trumbore against all triangles, up to 64 sorted crossings, inside/outside walk along the
 * column with the reference's error corrections. p0/p1/p2: 3 floats per triangle; bbu[16] as packed by LBM_Domain::voxelize_mesh_on_device (FX/lbm.cpp:529-549). */
void luwo_voxelize_mesh(const luwo_params* p, uint32_t direction, const float* u, uint8_t* flags, uint8_t flag, const float* p0, const float* p1, const float* p2, const float* bbu);

/* running mean / M2 of the sampled fields: the host loop `accumulate_from_buffers` of the case driver, FX/setup.cpp:4441-4488 (Welford update per cell,
 * u_avg interleaved [3n+c], products and sums rounded separately as the reference's g++ build does). `count` is the sample number AFTER the increment. */
void luwo_stats_accumulate(uint64_t N, uint64_t count, const float* rho, const float* u, float* u_avg, float* rho_avg, float* m2_u, float* m2_v, float* m2_w);

void luwo_set_threads(int n); /* 0 = all cores */
int luwo_get_threads(void);

#ifdef __cplusplus
}
#endif
#endif
