// C++ host layer of the B200 LBM (see lbm.hpp): domain split, initialisation and time-step orchestration of the reference's LBM object
// (FX/lbm.cpp:1057-1112, 1221-1312) on top of the C ABI. All device work is enqueued asynchronously by one host thread (FX/main.cpp:146-150).
#include "lbm.hpp"

LBM_Settings lbm_settings;
static uint env_uint(const char* name, const uint fallback) { const char* v = getenv(name); return v&&*v ? (uint)strtoul(v, nullptr, 10) : fallback; }

float lbm_kernel_literal(const float x) { // what `to_string(float)` + the OpenCL compiler make of a constant (FX/utilities.hpp:2741-2750): 8 decimals, round-trip through text
	if(!std::isfinite(x)) return x;
	float v = std::fabs(x);
	int e = 0;
	if(v>=10.0f) { const float lim[6] = {1e32f,1e16f,1e8f,1e4f,1e2f,1e1f}, mul[6] = {1e-32f,1e-16f,1e-8f,1e-4f,1e-2f,1e-1f}; const int de[6] = {32,16,8,4,2,1};
		for(int k=0; k<6; k++) if(v>=lim[k]) { v *= mul[k]; e += de[k]; } }
	if(v>0.0f&&v<=1.0f) { const float lim[6] = {1e-31f,1e-15f,1e-7f,1e-3f,1e-1f,1e0f}, mul[6] = {1e32f,1e16f,1e8f,1e4f,1e2f,1e1f}; const int de[6] = {32,16,8,4,2,1};
		for(int k=0; k<6; k++) if(v<lim[k]) { v *= mul[k]; e -= de[k]; } }
	unsigned long long integral = (unsigned long long)v;
	const float rem = (v-(float)integral)*1e8f;
	unsigned long long dec = (unsigned long long)rem;
	if(rem-(float)dec>=0.5f) { dec++; if(dec>=100000000ull) { dec = 0ull; integral++; if(integral>=10ull) { integral = 1ull; e++; } } }
	char text[64];
	if(e!=0) snprintf(text, sizeof(text), "%s%llu.%08lluE%d", x<0.0f ? "-" : "", integral, dec, e);
	else snprintf(text, sizeof(text), "%s%llu.%08llu", x<0.0f ? "-" : "", integral, dec);
	return strtof(text, nullptr);
}

#ifdef LUW_USE_REFERENCE_UTILITIES
// Inside the reference tree this file stands in for FX/lbm.cpp: it defines the globals that one defines (FX/lbm.cpp:18-20) and reads the ones the case driver
// sets before it constructs an LBM (FX/lbm.cpp:3-14; written by update_coriolis / update_buffer_nudging / update_top_sponge, FX/setup.cpp:3800-3903), which the
// reference bakes into its kernel source (FX/lbm.cpp:770-782) -- so the case driver needs no edits for them.
Units units; // for unit conversion
float coriolis_f_lbmu = 0.0f;
extern bool buffer_nudging_active; extern int buffer_n_cells; extern float buffer_inv_tau_lbmu; extern int buffer_nudge_vertical; extern int buffer_downstream_face_id;
extern bool top_sponge_active; extern int sponge_n_cells; extern float sponge_inv_tau_lbmu; extern int sponge_ref_mode;
static void settings_from_case_driver() {
	lbm_settings.downstream_face = buffer_downstream_face_id; // def_downstream_face
	lbm_settings.set_buffer_nudging(buffer_nudging_active&&buffer_n_cells>0 ? (uint)buffer_n_cells : 0u, buffer_inv_tau_lbmu, buffer_nudge_vertical!=0);
	lbm_settings.set_top_sponge(top_sponge_active&&sponge_n_cells>0 ? (uint)sponge_n_cells : 0u, sponge_inv_tau_lbmu);
	if(top_sponge_active&&sponge_ref_mode!=0) print_error("sponge_ref_mode "+to_string(sponge_ref_mode)+" is not implemented (the reference kernel implements mode 0 only, FX/kernel.cpp:1597).");
#ifdef TEMPERATURE // the reference's compile-time switch (FX/defines.hpp:23) left on in the tree this is built in: the case driver then writes lbm.T and passes alpha / beta
	lbm_settings.features |= LUW_TEMPERATURE;
#endif
}
float3 vtk_origin_shift = float3(0.0f, 0.0f, 0.0f); // FX/lbm.cpp:18-20
string default_filename(const string& path, const string& name, const string& extension, const ulong t) { // FX/lbm.cpp:235-239: <path or exe/export/><name>-<9-digit step><extension>
	string time = "00000000"+to_string(t);
	time = substring(time, length(time)-9u, 9u);
	return (path=="" ? get_exe_path()+"export/" : path)+create_file_extension((name=="" ? "file" : name)+"-"+time, extension);
}
string default_filename(const string& name, const string& extension, const ulong t) { return default_filename("", name, extension, t); }
// FX/lbm.cpp:95-142 for this build: host mirrors rho, u, flags; device the same + 19 DDFs; traffic per step = 19 loads + 19 stores + flags (+ rho, u stores with UPDATE_FIELDS)
static uint ddf_bytes() { return env_uint("LUW_PRECISION", lbm_settings.precision)==LUW_FP32 ? 4u : 2u; }
static bool thermal_on() { return (lbm_settings.features&LUW_TEMPERATURE)!=0u; }
uint bytes_per_cell_host() { return 17u+(thermal_on() ? 4u : 0u); }
uint bytes_per_cell_device() { return 19u*ddf_bytes()+17u+(thermal_on() ? 7u*ddf_bytes()+4u : 0u); }
uint bandwidth_bytes_per_cell_device() { return 38u*ddf_bytes()+1u+((lbm_settings.features&LUW_UPDATE_FIELDS) ? 16u : 0u)+(thermal_on() ? 14u*ddf_bytes()+((lbm_settings.features&LUW_UPDATE_FIELDS) ? 4u : 0u) : 0u); } // + 7 g loads + 7 g stores (+ T), FX/lbm.cpp:127-128
#endif // LUW_USE_REFERENCE_UTILITIES
// Device memory of THIS build per domain: DDFs (19 fpxx) + rho, u (16 B) + flags (1 B) per cell of the padded local lattice, + halo buffers. The reference's
// estimator (FX/lbm.cpp:143-232) counts its own buffer set (F, gi, T, graphics, transfer buffers): a deck with mesh_control="gpu_memory" therefore resolves to a
// finer grid here than there for the same number of MB -- use mesh_control="cell_size" where the grid must be identical.
uint vram_required_mb_per_device(const uint Nx, const uint Ny, const uint Nz, const uint Dx, const uint Dy, const uint Dz) {
	const ulong lx = (ulong)(Nx/Dx+2u*(Dx>1u)), ly = (ulong)(Ny/Dy+2u*(Dy>1u)), lz = (ulong)(Nz/Dz+2u*(Dz>1u));
	const ulong px = (lx+15ull)&~15ull;
	const ulong ddf = (env_uint("LUW_PRECISION", lbm_settings.precision)==LUW_FP32) ? 4ull : 2ull;
	const ulong cells = px*ly*lz;
	const ulong halo = 8ull*(ddf==4ull ? 20ull : 17ull)*((Dx>1u ? ly*lz : 0ull)+(Dy>1u ? lz*lx : 0ull)+(Dz>1u ? lx*ly : 0ull));
	const ulong thermal = (lbm_settings.features&LUW_TEMPERATURE) ? 7ull*ddf+4ull+12ull : 0ull; // gi + T (FX/lbm.cpp:124-126) + the pre-force velocity of the two-kernel thermal step
	return (uint)((cells*(19ull*ddf+17ull+thermal)+halo)/1048576ull)+1u;
}
uint vram_required_mb_total(const uint Nx, const uint Ny, const uint Nz, const uint Dx, const uint Dy, const uint Dz) { return Dx*Dy*Dz*vram_required_mb_per_device(Nx, Ny, Nz, Dx, Dy, Dz); }

uint LBM_Domain::lbm_features() { return lbm_settings.features; }

// FX/lbm.cpp:52-90 (triangle_aabb, make_projected_bounds, overlap_1d, overlaps_projected) and :1457-1483 (collect_domain_triangle_ids), pad 1 cell, eps 1e-4
std::vector<uint> luw_cull_triangles(const float* p0, const float* p1, const float* p2, const uint triangle_number, const uint direction,
	const int Ox, const int Oy, const int Oz, const uint local_Nx, const uint local_Ny, const uint local_Nz) {
	const float lo[3] = { (float)Ox, (float)Oy, (float)Oz };
	const float hi[3] = { (float)(Ox+(int)local_Nx-1), (float)(Oy+(int)local_Ny-1), (float)(Oz+(int)local_Nz-1) };
	const uint a = direction==0u ? 1u : 0u, b = direction==2u ? 1u : 2u; // the two axes of the projection plane: yz, xz, xy
	const float pad = 1.0f, eps = 1.0e-4f;
	std::vector<uint> ids;
	ids.reserve(triangle_number/4u+1u);
	for(uint i=0u; i<triangle_number; i++) {
		bool keep = true;
		for(const uint axis : { a, b }) {
			const float v0 = p0[3u*i+axis], v1 = p1[3u*i+axis], v2 = p2[3u*i+axis];
			const float tmin = fmin(fmin(v0, v1), v2), tmax = fmax(fmax(v0, v1), v2);
			keep = keep&&(tmax>=lo[axis]-pad-eps&&tmin<=hi[axis]+pad+eps);
		}
		if(keep) ids.push_back(i);
	}
	return ids;
}
// test hook (tests/test_voxelize.py, no device needed): ids_out has room for triangle_number entries; returns the count
extern "C" uint luw_host_cull_triangles(const float* p0, const float* p1, const float* p2, const uint triangle_number, const uint direction,
	const int Ox, const int Oy, const int Oz, const uint local_Nx, const uint local_Ny, const uint local_Nz, uint* ids_out) {
	const std::vector<uint> ids = luw_cull_triangles(p0, p1, p2, triangle_number, direction, Ox, Oy, Oz, local_Nx, local_Ny, local_Nz);
	for(size_t i=0u; i<ids.size(); i++) ids_out[i] = ids[i];
	return (uint)ids.size();
}

// ---------------------------------------------------------------------------------------------------------------- test hook: LUW_DUMP_DIR
static const char* dump_dir() { static const char* d = getenv("LUW_DUMP_DIR"); return (d&&d[0]) ? d : nullptr; }
bool luw_dump_enabled() { return dump_dir()!=nullptr; }
void luw_dump_array(const char* name, const void* data, const ulong bytes) {
	const std::string path = std::string(dump_dir())+"/"+name+".bin";
	FILE* f = fopen(path.c_str(), "wb");
	if(!f) return;
	fwrite(data, 1u, (size_t)bytes, f);
	fclose(f);
}
void luw_dump_voxelize_call(luw_domain* handle, const ulong N, const uint direction, const uchar flag, const float* p0, const float* p1, const float* p2, const uint triangle_number, const float* bbu) {
	const uint head[4] = { triangle_number, direction, (uint)flag, 0u };
	std::vector<float> tri; tri.reserve(16u+9u*(size_t)triangle_number);
	tri.insert(tri.end(), bbu, bbu+16);
	tri.insert(tri.end(), p0, p0+3u*(size_t)triangle_number); tri.insert(tri.end(), p1, p1+3u*(size_t)triangle_number); tri.insert(tri.end(), p2, p2+3u*(size_t)triangle_number);
	luw_dump_array("vox_head", head, sizeof(head));
	luw_dump_array("vox_tri", tri.data(), tri.size()*sizeof(float));
	std::vector<uchar> fl((size_t)N); std::vector<float> u(3u*(size_t)N); // the device images the kernel works on
	luw_check(luw_download(handle, LUW_FIELD_FLAGS, fl.data(), 0ull, N)); luw_check(luw_download(handle, LUW_FIELD_U, u.data(), 0ull, 3ull*N)); luw_check(luw_sync(handle));
	luw_dump_array("vox_flags_before", fl.data(), N); luw_dump_array("vox_u_before", u.data(), 3ull*N*sizeof(float));
}
void LBM::dump_state(const char* tag) {
	if(!luw_dump_enabled()||get_D()!=1u) return;
	LBM_Domain* dm = lbm_domain[0];
	const ulong N = dm->get_N();
	const std::string t(tag);
	if(t=="init") { // the host images initialize() uploads, and the constants the kernels were given
		luw_dump_array("init_flags", dm->flags.data(), N); luw_dump_array("init_rho", dm->rho.data(), N*sizeof(float)); luw_dump_array("init_u", dm->u.data(), 3ull*N*sizeof(float));
		const std::string path = std::string(dump_dir())+"/params.txt";
		FILE* f = fopen(path.c_str(), "w");
		if(f) {
			fprintf(f, "Nx %u\nNy %u\nNz %u\nprecision %u\nfeatures %u\narith %u\n", dm->get_Nx(), dm->get_Ny(), dm->get_Nz(), env_uint("LUW_PRECISION", lbm_settings.precision), lbm_settings.features, env_uint("LUW_ARITH", lbm_settings.arith));
			fprintf(f, "w %.9g\nfx %.9g\nfy %.9g\nfz %.9g\nomega_x %.9g\nomega_y %.9g\nomega_z %.9g\n", (double)lbm_kernel_literal(1.0f/(3.0f*dm->get_nu()+0.5f)), (double)dm->get_fx(), (double)dm->get_fy(), (double)dm->get_fz(), (double)dm->get_omega_x(), (double)dm->get_omega_y(), (double)dm->get_omega_z());
			fprintf(f, "downstream_face %d\nbuffer_N %u\nbuffer_inv_tau %.9g\nbuffer_nudge_vertical %d\nsponge_N %u\nsponge_inv_tau %.9g\n", lbm_settings.downstream_face, lbm_settings.buffer_N>0u ? lbm_settings.buffer_N : 1u, (double)lbm_kernel_literal(lbm_settings.buffer_inv_tau),
				lbm_settings.buffer_nudge_vertical, lbm_settings.sponge_N>0u ? lbm_settings.sponge_N : 1u, (double)lbm_kernel_literal(lbm_settings.sponge_inv_tau));
			fclose(f);
		}
		return;
	}
	std::vector<float> r((size_t)N), u(3u*(size_t)N); // rho / u of the device, without touching the host mirrors the case driver works on
	if(!(lbm_settings.features&LUW_UPDATE_FIELDS)) dm->enqueue_update_fields();
	luw_check(luw_download(dm->get_handle(), LUW_FIELD_RHO, r.data(), 0ull, N)); luw_check(luw_download(dm->get_handle(), LUW_FIELD_U, u.data(), 0ull, 3ull*N)); luw_check(luw_sync(dm->get_handle()));
	luw_dump_array((t+"_rho").c_str(), r.data(), N*sizeof(float)); luw_dump_array((t+"_u").c_str(), u.data(), 3ull*N*sizeof(float));
}

luw_domain* LBM_Domain::create_handle(const int device, const uint Nx, const uint Ny, const uint Nz, const uint Dx, const uint Dy, const uint Dz, const int Ox, const int Oy, const int Oz, const float nu, const float alpha, const float beta) {
	luw_domain_params p;
	memset(&p, 0, sizeof(p));
	p.Nx = Nx; p.Ny = Ny; p.Nz = Nz; p.Dx = Dx; p.Dy = Dy; p.Dz = Dz; p.Ox = Ox; p.Oy = Oy; p.Oz = Oz;
	p.precision = env_uint("LUW_PRECISION", lbm_settings.precision);
	p.features = lbm_settings.features;
	p.arith = env_uint("LUW_ARITH", lbm_settings.arith);
	p.w = lbm_kernel_literal(1.0f/(3.0f*nu+0.5f)); // def_w, FX/lbm.hpp:146 + FX/lbm.cpp:663
	p.downstream_face = lbm_settings.downstream_face;
	p.buffer_N = lbm_settings.buffer_N>0u ? lbm_settings.buffer_N : 1u; p.buffer_inv_tau = lbm_kernel_literal(lbm_settings.buffer_inv_tau); p.buffer_nudge_vertical = lbm_settings.buffer_nudge_vertical;
	p.sponge_N = lbm_settings.sponge_N>0u ? lbm_settings.sponge_N : 1u; p.sponge_inv_tau = lbm_kernel_literal(lbm_settings.sponge_inv_tau);
	p.device = device;
	luw_domain* h = nullptr;
	luw_check(luw_domain_create(&p, &h));
	if(p.features&LUW_TEMPERATURE) luw_check(luw_thermal_params(h, lbm_kernel_literal(1.0f/(2.0f*alpha+0.5f)), lbm_kernel_literal(beta), lbm_kernel_literal(1.0f))); // def_w_T, def_beta, def_T_avg: FX/lbm.cpp:750-752
	return h;
}

LBM_Domain::LBM_Domain(const int device, const uint Nx, const uint Ny, const uint Nz, const uint Dx, const uint Dy, const uint Dz, const int Ox, const int Oy, const int Oz,
	const float nu, const float fx, const float fy, const float fz, const float alpha, const float beta)
	: Nx(Nx), Ny(Ny), Nz(Nz), Dx(Dx), Dy(Dy), Dz(Dz), Ox(Ox), Oy(Oy), Oz(Oz), nu(nu), fx(fx), fy(fy), fz(fz), alpha(alpha), beta(beta), device(device),
	  handle(create_handle(device, Nx, Ny, Nz, Dx, Dy, Dz, Ox, Oy, Oz, nu, alpha, beta)),
	  rho(handle, LUW_FIELD_RHO, (ulong)Nx*Ny*Nz, 1u, 1.0f), u(handle, LUW_FIELD_U, (ulong)Nx*Ny*Nz, 3u, 0.0f), flags(handle, LUW_FIELD_FLAGS, (ulong)Nx*Ny*Nz, 1u, (uchar)0) {
	if(thermal()) T = Memory<float>(handle, LUW_FIELD_T, (ulong)Nx*Ny*Nz, 1u, 1.0f); // FX/lbm.cpp:323
}

LBM_Domain::~LBM_Domain() { luw_domain_destroy(handle); }

// ---------------------------------------------------------------------------------------------------------------- LBM
void LBM::construct(const uint Nx_, const uint Ny_, const uint Nz_, const uint Dx_, const uint Dy_, const uint Dz_, const float nu, const float fx, const float fy, const float fz, const float alpha, const float beta) {
#ifdef LUW_USE_REFERENCE_UTILITIES
	settings_from_case_driver();
#endif
	if(env_uint("LUW_TEMPERATURE", (lbm_settings.features&LUW_TEMPERATURE) ? 1u : 0u)) lbm_settings.features |= LUW_TEMPERATURE; else lbm_settings.features &= ~(uint)LUW_TEMPERATURE;
	// Without LUW_TEMPERATURE alpha / beta are ignored, NOT an error as in FX/lbm.cpp:1162: the case driver passes units.alpha(si_alpha_air) to every LBM it builds
	// (FX/setup.cpp:3740, 4935, 5720, 6018) whatever defines.hpp says, and the drop-in driver is built with TEMPERATURE off by default.
	if(Dx_*Dy_*Dz_==0u) print_error("You specified 0 LBM grid domains. There has to be at least 1 domain in every direction.");
	Dx = Dx_; Dy = Dy_; Dz = Dz_;
	Nx = (Nx_/Dx)*Dx; Ny = (Ny_/Dy)*Dy; Nz = (Nz_/Dz)*Dz; // global size rounded down to multiples of the domain counts, FX/lbm.cpp:1058-1060
	if(Nx*Ny*Nz==0u) print_error("Grid point number is 0.");
	const uint D = Dx*Dy*Dz;
	int ndev = 0;
	luw_check(luw_device_count(&ndev));
	if(ndev<1) print_error("No CUDA device is available; this build of the LBM has no CPU fallback.");
	const uint Hx = Dx>1u, Hy = Dy>1u, Hz = Dz>1u; // halo offsets, FX/lbm.cpp:1062-1064
	// Device assignment, smart_device_selection of FX/lbm.cpp:947-1034: D cards of ONE model, the fastest model that has D of them, never two domains on one card --
	// except that a box with fewer cards than domains is not an error here but a warning (several domains then share a device round robin; the parity tests run
	// 2 x 2 x 2 decompositions on one GPU that way; LUW_STRICT_DEVICES=1 restores the reference's error). lbm_settings.devices overrides the selection.
	std::vector<int> chosen(D, 0);
	{
		std::vector<luw_device_info> info((size_t)ndev);
		for(int i=0; i<ndev; i++) luw_check(luw_get_device_info(i, &info[(size_t)i]));
		if((int)D>ndev) {
			const std::string msg = "Domain partition count ("+std::to_string(D)+") exceeds available physical cards ("+std::to_string((uint)ndev)+").";
			if(env_uint("LUW_STRICT_DEVICES", 0u)) print_error(msg+" Reduce Dx*Dy*Dz to <= "+std::to_string((uint)ndev)+".");
			fprintf(stderr, "[luw] warning: %s Domains share devices round robin.\n", msg.c_str());
			for(uint d=0u; d<D; d++) chosen[d] = (int)(d%(uint)ndev);
		} else {
			int best = -1; double best_value = -1.0; // "tflops" of a card: SMs x clock (all models here have the same lanes per SM)
			for(int i=0; i<ndev; i++) {
				int same = 0;
				for(int k=0; k<ndev; k++) same += std::string(info[(size_t)k].name)==std::string(info[(size_t)i].name);
				const double value = (double)info[(size_t)i].compute_units*(double)info[(size_t)i].clock_mhz;
				if(same>=(int)D&&value>best_value) { best_value = value; best = i; }
			}
			uint d = 0u;
			if(best>=0) { for(int i=0; i<ndev&&d<D; i++) if(std::string(info[(size_t)i].name)==std::string(info[(size_t)best].name)) chosen[d++] = i; }
			else for(int i=0; i<ndev&&d<D; i++) chosen[d++] = i; // mixed models without oversubscription
		}
		// VRAM preflight, sanity_checks_constructor of FX/lbm.cpp:1123-1140: the message of the reference, before anything is allocated
		uint memory_available = 0xFFFFFFFFu;
		std::vector<uint> share((size_t)ndev, 0u);
		for(uint d=0u; d<D; d++) { const int dev = d<lbm_settings.devices.size() ? lbm_settings.devices[d] : chosen[d]; if(dev>=0&&dev<ndev) share[(size_t)dev]++; }
		for(int i=0; i<ndev; i++) if(share[(size_t)i]>0u) memory_available = std::min(memory_available, (uint)(info[(size_t)i].memory_bytes/1048576ull)/share[(size_t)i]);
		const uint memory_required = vram_required_mb_per_device(Nx, Ny, Nz, Dx, Dy, Dz);
		if(memory_required>memory_available) {
			const float factor = cbrtf((float)memory_available/(float)memory_required);
			print_error("Grid resolution ("+std::to_string(Nx)+", "+std::to_string(Ny)+", "+std::to_string(Nz)+") is too large: "+std::to_string(D)+"x "+std::to_string(memory_required)+" MB required, "+std::to_string(D)+"x "+std::to_string(memory_available)
				+" MB available. Largest possible resolution is ("+std::to_string((uint)(factor*(float)Nx))+", "+std::to_string((uint)(factor*(float)Ny))+", "+std::to_string((uint)(factor*(float)Nz))+"). Restart the simulation with lower resolution or on different device(s) with more memory.");
		}
	}
	lbm_domain = new LBM_Domain*[D];
	for(uint d=0u; d<D; d++) {
		const uint x = (d%(Dx*Dy))%Dx, y = (d%(Dx*Dy))/Dx, z = d/(Dx*Dy); // d = x+(y+z*Dy)*Dx, FX/lbm.cpp:1066-1073
		const int device = d<lbm_settings.devices.size() ? lbm_settings.devices[d] : chosen[d];
		lbm_domain[d] = new LBM_Domain(device, Nx/Dx+2u*Hx, Ny/Dy+2u*Hy, Nz/Dz+2u*Hz, Dx, Dy, Dz, (int)(x*Nx/Dx)-(int)Hx, (int)(y*Ny/Dy)-(int)Hy, (int)(z*Nz/Dz)-(int)Hz, nu, fx, fy, fz, alpha, beta);
		handles.push_back(lbm_domain[d]->get_handle());
		if(LBM_Domain::thermal()) T_buffers.push_back(&lbm_domain[d]->T);
		rho_buffers.push_back(&lbm_domain[d]->rho); u_buffers.push_back(&lbm_domain[d]->u); flags_buffers.push_back(&lbm_domain[d]->flags);
	}
	rho.bind(this, rho_buffers.data(), "rho"); u.bind(this, u_buffers.data(), "u"); flags.bind(this, flags_buffers.data(), "flags");
	if(LBM_Domain::thermal()) T.bind(this, T_buffers.data(), "T");
}
LBM::LBM(const uint Nx, const uint Ny, const uint Nz, const uint Dx, const uint Dy, const uint Dz, const float nu, const float fx, const float fy, const float fz, const float, const float alpha, const float beta) { construct(Nx, Ny, Nz, Dx, Dy, Dz, nu, fx, fy, fz, alpha, beta); }
LBM::LBM(const uint Nx, const uint Ny, const uint Nz, const float nu, const float fx, const float fy, const float fz, const float, const float alpha, const float beta) { construct(Nx, Ny, Nz, 1u, 1u, 1u, nu, fx, fy, fz, alpha, beta); }
LBM::LBM(const uint3 N, const uint Dx, const uint Dy, const uint Dz, const float nu, const float fx, const float fy, const float fz, const float, const float alpha, const float beta) { construct(N.x, N.y, N.z, Dx, Dy, Dz, nu, fx, fy, fz, alpha, beta); }
LBM::LBM(const uint3 N, const float nu, const float fx, const float fy, const float fz, const float, const float alpha, const float beta) { construct(N.x, N.y, N.z, 1u, 1u, 1u, nu, fx, fy, fz, alpha, beta); }
LBM::~LBM() {
#ifdef LUW_USE_REFERENCE_UTILITIES
	flush_pending();
	info.print_finalize(); // FX/lbm.cpp: the console table is closed with the simulation
#endif
	for(uint d=0u; d<get_D(); d++) delete lbm_domain[d];
	delete[] lbm_domain;
}

void LBM::communicate(const int payload) { // FX/lbm.cpp:1907-1958: x, then y, then z
	if(get_D()==1u) return;
	for(uint axis=0u; axis<3u; axis++) luw_check(luw_halo_exchange(handles.data(), get_D(), payload, axis, lbm_domain[0]->get_t()));
}
void LBM::initialize() { // FX/lbm.cpp:1221-1260
	const uint D = get_D();
	dump_state("init");
	for(uint d=0u; d<D; d++) { lbm_domain[d]->rho.enqueue_write_to_device(); lbm_domain[d]->u.enqueue_write_to_device(); lbm_domain[d]->flags.enqueue_write_to_device(); }
	if(LBM_Domain::thermal()) for(uint d=0u; d<D; d++) lbm_domain[d]->T.enqueue_write_to_device();
	for(uint d=0u; d<D; d++) lbm_domain[d]->increment_time_step(); // slot parity t = 1 for the initial DDF layout
	communicate(LUW_HALO_RHO_U_FLAGS);
	for(uint d=0u; d<D; d++) lbm_domain[d]->enqueue_initialize();
	communicate(LUW_HALO_RHO_U_FLAGS);
	communicate(LUW_HALO_FI);
	if(LBM_Domain::thermal()) { communicate(LUW_HALO_T); communicate(LUW_HALO_GI); } // communicate_T(); communicate_gi(); (time step must be odd here)
	for(uint d=0u; d<D; d++) lbm_domain[d]->finish_queue();
	for(uint d=0u; d<D; d++) lbm_domain[d]->reset_time_step();
	initialized = true;
}
void LBM::do_time_step() { // FX/lbm.cpp:1262-1290 (GRAPHICS exchanges and the per-step finish_queue are gone)
	const uint D = get_D();
	for(uint d=0u; d<D; d++) lbm_domain[d]->enqueue_stream_collide();
	communicate(LUW_HALO_FI);
	if(LBM_Domain::thermal()) communicate(LUW_HALO_GI); // communicate_gi()
	for(uint d=0u; d<D; d++) lbm_domain[d]->increment_time_step();
	static const ulong dump_step = []{ const char* e = getenv("LUW_DUMP_STEP"); return e ? (ulong)atol(e) : 0ull; }();
	if(dump_step>0ull&&get_t()==dump_step) dump_state("step");
}
void LBM::run(const ulong steps, const ulong total_steps) { // FX/lbm.cpp:1292-1312
#ifdef LUW_USE_REFERENCE_UTILITIES // inside the reference tree: feed the console table / ETA model exactly like FX/lbm.cpp does (FX/info.cpp reads it from another thread)
	info.append(steps, total_steps, get_t());
	if(!initialized) { initialize(); info.print_initialize(this); }
	// The case driver calls run(1) once per step (FX/setup.cpp:4898). Synchronising after every step, as FX/lbm.cpp:1306 does, leaves the GPU idle while the host
	// prepares the next step (14.8 GLUP/s on the console for the 3.7 M-cell example deck); here the steps stay enqueued and the queues are drained every
	// LUW_SYNC_EVERY steps (default 16; 1 = the reference's behaviour). The elapsed time is credited to the drained steps in equal parts, so `info` still sees one
	// update per step. Host reads in between (probes, averages, VTK) synchronise by themselves: they are ordered on the same streams.
	static const ulong sync_every = []{ const char* e = getenv("LUW_SYNC_EVERY"); const long v = e ? atol(e) : 16l; return (ulong)(v<1l ? 1l : v); }();
	const ulong n = steps==max_ulong ? 0ull : steps;
	for(ulong i=0ull; i<n; i++) {
		if(pending_steps==0ull) pending_t0 = std::chrono::high_resolution_clock::now();
		do_time_step();
		if(++pending_steps>=sync_every) flush_pending();
	}
#else // stand-alone: steps are enqueued back to back, one synchronisation at the end of the call
	(void)total_steps;
	if(!initialized) initialize();
	const ulong n = steps==max_ulong ? 0ull : steps; // the reference's "run forever" needs the interactive main loop; not part of this layer
	for(ulong i=0ull; i<n; i++) do_time_step();
	for(uint d=0u; d<get_D(); d++) lbm_domain[d]->finish_queue();
#endif
}
void LBM::flush_pending() {
	if(pending_steps==0ull) return;
	for(uint d=0u; d<get_D(); d++) lbm_domain[d]->finish_queue();
#ifdef LUW_USE_REFERENCE_UTILITIES
	const double dt = std::chrono::duration_cast<std::chrono::duration<double>>(std::chrono::high_resolution_clock::now()-pending_t0).count()/(double)pending_steps;
	for(ulong k=0ull; k<pending_steps; k++) info.update(dt);
#endif
	pending_steps = 0ull;
}
void LBM::update_fields() { for(uint d=0u; d<get_D(); d++) lbm_domain[d]->enqueue_update_fields(); for(uint d=0u; d<get_D(); d++) lbm_domain[d]->finish_queue(); }
void LBM::reset() { initialized = false; }
