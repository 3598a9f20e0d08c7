// TEST INFRASTRUCTURE (oracle/): host-thread "launcher" for the reference's own kernel text.
// Build (see oracle/Makefile):  g++ -O2 -ffp-contract=off -fopenmp -D<features> -D<FP16S|FP16C> ref_unit.cpp -> oracle/_ref/libluwref_<variant>.so
// Each exported function mirrors one enqueue of the reference host code: a 1-D NDRange over N cells (or A face cells,
// or P inlet points) with one work-item per index (FX/opencl.hpp:649-670). Work-items are independent by construction
// of the in-place Esoteric-Pull layout (FX/kernel.cpp:1338-1351), so running them on OpenMP threads in any order
// reproduces what an OpenCL device computes.
#include "clshim.hpp"

namespace refcl {
#include "../_ref/kernel_cl.inc"
} // namespace refcl

extern "C" {

struct luwref_params { // one-to-one with the def_* constants of FX/lbm.cpp:612-783 that the LBM kernels read
	uint32_t Nx, Ny, Nz; // local lattice incl. halo layers
	uint32_t Dx, Dy, Dz; // number of domains per axis
	int32_t Ox, Oy, Oz; // global coordinate of local cell 0
	float w; // def_w = 1/tau
	int32_t downstream_face; // 0 none, 1 west, 2 east, 3 south, 4 north
	uint32_t buffer_N; float buffer_inv_tau; int32_t buffer_nudge_vertical;
	uint32_t sponge_N; float sponge_inv_tau;
};

void luwref_set_params(const luwref_params* p) {
	g_Nx = p->Nx; g_Ny = p->Ny; g_Nz = p->Nz; g_N = (ulong)p->Nx*(ulong)p->Ny*(ulong)p->Nz;
	g_Dx = p->Dx; g_Dy = p->Dy; g_Dz = p->Dz; g_Ox = p->Ox; g_Oy = p->Oy; g_Oz = p->Oz;
	g_Nx_global = (g_Nx-2u*(g_Dx>1u))*g_Dx; // same derivations as LBM_Domain::device_defines()
	g_Ny_global = (g_Ny-2u*(g_Dy>1u))*g_Dy;
	g_Nz_global = (g_Nz-2u*(g_Dz>1u))*g_Dz;
	g_west_local_x = -g_Ox; g_east_local_x = (int)g_Nx_global-1-g_Ox;
	g_south_local_y = -g_Oy; g_north_local_y = (int)g_Ny_global-1-g_Oy;
	g_top_local_z = (int)g_Nz_global-1-g_Oz;
	g_has_west = g_west_local_x>=0&&g_west_local_x<(int)g_Nx;
	g_has_east = g_east_local_x>=0&&g_east_local_x<(int)g_Nx;
	g_has_south = g_south_local_y>=0&&g_south_local_y<(int)g_Ny;
	g_has_north = g_north_local_y>=0&&g_north_local_y<(int)g_Ny;
	g_has_top = g_top_local_z>=0&&g_top_local_z<(int)g_Nz;
	g_w = p->w;
	g_downstream_face = p->downstream_face;
	g_buffer_N = p->buffer_N; g_buffer_inv_tau = p->buffer_inv_tau; g_buffer_nudge_vertical = p->buffer_nudge_vertical;
	g_sponge_N = p->sponge_N; g_sponge_inv_tau = p->sponge_inv_tau;
}

uint32_t luwref_sizeof_fpxx() { return (uint32_t)sizeof(fpxx); }
uint32_t luwref_features() { // which compile-time switches this shared object was built with
	uint32_t f = 0u;
#ifdef UPDATE_FIELDS
	f |= 1u;
#endif
#ifdef VOLUME_FORCE
	f |= 2u;
#endif
#ifdef EQUILIBRIUM_BOUNDARIES
	f |= 4u;
#endif
#ifdef SUBGRID
	f |= 8u;
#endif
#ifdef BUFFER_NUDGING
	f |= 16u;
#endif
#ifdef TOP_SPONGE
	f |= 32u;
#endif
#ifdef TEMPERATURE
	f |= 64u;
#endif
	return f;
}

#define NDRANGE(count, call) { const long long cnt_ = (long long)(count); _Pragma("omp parallel for schedule(static)") for(long long i_=0; i_<cnt_; i_++) { cl_gid = (ulong)i_; call; } }

#ifndef TEMPERATURE
void luwref_initialize(fpxx* fi, const float* rho, float* u, uchar* flags) {
	NDRANGE(g_N, refcl::initialize(fi, rho, u, flags));
}
void luwref_stream_collide(fpxx* fi, float* rho, float* u, uchar* flags, const ulong t, const float fx, const float fy, const float fz, const float ox, const float oy, const float oz) {
	NDRANGE(g_N, refcl::stream_collide(fi, rho, u, flags, t, fx, fy, fz, ox, oy, oz));
}
void luwref_update_fields(const fpxx* fi, float* rho, float* u, const uchar* flags, const ulong t, const float fx, const float fy, const float fz, const float ox, const float oy, const float oz) {
	NDRANGE(g_N, refcl::update_fields(fi, rho, u, flags, t, fx, fy, fz, ox, oy, oz));
}
#else // TEMPERATURE: the same kernels take the D3Q7 DDFs and the temperature field as trailing arguments (FX/kernel.cpp:1374-1376, 1482-1484, 1942-1944)
void luwref_set_thermal(const float w_T, const float beta, const float T_avg) { g_w_T = w_T; g_beta = beta; g_T_avg = T_avg; } // def_w_T, def_beta, def_T_avg (FX/lbm.cpp:750-752)
void luwref_initialize_thermal(fpxx* fi, const float* rho, float* u, uchar* flags, fpxx* gi, const float* T) {
	NDRANGE(g_N, refcl::initialize(fi, rho, u, flags, gi, T));
}
void luwref_stream_collide_thermal(fpxx* fi, float* rho, float* u, uchar* flags, const ulong t, const float fx, const float fy, const float fz, const float ox, const float oy, const float oz, fpxx* gi, float* T) {
	NDRANGE(g_N, refcl::stream_collide(fi, rho, u, flags, t, fx, fy, fz, ox, oy, oz, gi, T));
}
void luwref_update_fields_thermal(const fpxx* fi, float* rho, float* u, const uchar* flags, const ulong t, const float fx, const float fy, const float fz, const float ox, const float oy, const float oz, const fpxx* gi, float* T) {
	NDRANGE(g_N, refcl::update_fields(fi, rho, u, flags, t, fx, fy, fz, ox, oy, oz, gi, T));
}
#endif
static ulong area(const uint direction) { const ulong A[3] = {(ulong)g_Ny*g_Nz, (ulong)g_Nz*g_Nx, (ulong)g_Nx*g_Ny}; return A[direction]; }
void luwref_transfer_extract_fi(const uint direction, const ulong t, fpxx_copy* bp, fpxx_copy* bm, const fpxx_copy* fi) {
	NDRANGE(area(direction), refcl::transfer_extract_fi(direction, t, bp, bm, fi));
}
void luwref_transfer_insert_fi(const uint direction, const ulong t, const fpxx_copy* bp, const fpxx_copy* bm, fpxx_copy* fi) {
	NDRANGE(area(direction), refcl::transfer__insert_fi(direction, t, bp, bm, fi));
}
void luwref_transfer_extract_rho_u_flags(const uint direction, const ulong t, char* bp, char* bm, const float* rho, const float* u, const uchar* flags) {
	NDRANGE(area(direction), refcl::transfer_extract_rho_u_flags(direction, t, bp, bm, rho, u, flags));
}
void luwref_transfer_insert_rho_u_flags(const uint direction, const ulong t, const char* bp, const char* bm, float* rho, float* u, uchar* flags) {
	NDRANGE(area(direction), refcl::transfer__insert_rho_u_flags(direction, t, bp, bm, rho, u, flags));
}
#ifdef TEMPERATURE
void luwref_transfer_extract_gi(const uint direction, const ulong t, fpxx_copy* bp, fpxx_copy* bm, const fpxx_copy* gi) {
	NDRANGE(area(direction), refcl::transfer_extract_gi(direction, t, bp, bm, gi));
}
void luwref_transfer_insert_gi(const uint direction, const ulong t, const fpxx_copy* bp, const fpxx_copy* bm, fpxx_copy* gi) {
	NDRANGE(area(direction), refcl::transfer__insert_gi(direction, t, bp, bm, gi));
}
void luwref_transfer_extract_T(const uint direction, const ulong t, float* bp, float* bm, const float* T) {
	NDRANGE(area(direction), refcl::transfer_extract_T(direction, t, bp, bm, T));
}
void luwref_transfer_insert_T(const uint direction, const ulong t, const float* bp, const float* bm, float* T) {
	NDRANGE(area(direction), refcl::transfer__insert_T(direction, t, bp, bm, T));
}
#endif
void luwref_vk_inlet_apply(const uint use_interp, const float t0, const float t1, const float alpha, const ulong point_count, const ulong mode_count, const ulong mode_stride,
	const ulong* point_cell, const uchar* point_face, const float* point_data, const float* mode_data, float* u) {
	NDRANGE(point_count, refcl::vk_inlet_apply(use_interp, t0, t1, alpha, point_count, mode_count, mode_stride, point_cell, point_face, point_data, mode_data, u));
}
void luwref_voxelize_mesh(const uint direction, fpxx* fi, float* u, uchar* flags, const ulong t, const uchar flag, const float* p0, const float* p1, const float* p2, const float* bbu) {
	NDRANGE(area(direction), refcl::voxelize_mesh(direction, fi, u, flags, t, flag, p0, p1, p2, bbu));
}
float luwref_half_to_float_custom(const ushort x) { return refcl::half_to_float_custom(x); }
ushort luwref_float_to_half_custom(const float x) { return refcl::float_to_half_custom(x); }
void luwref_calculate_f_eq(const float rho, const float ux, const float uy, const float uz, float* feq) { refcl::calculate_f_eq(rho, ux, uy, uz, feq); }

} // extern "C"
