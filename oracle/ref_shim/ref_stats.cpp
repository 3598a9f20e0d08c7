// TEST INFRASTRUCTURE (oracle/): host shim around the reference's own averaging loop (FX/setup.cpp:4441-4488), whose text make_ref_stats.py extracts
// into oracle/_ref/stats_lambda.inc. Built like the reference builds it (g++ -O, no -march: products and sums are rounded separately).
#include <cstdint>
#include <vector>
typedef uint64_t ulong_t;
#define ulong ulong_t
struct Comp { const float* p; float operator[](ulong n) const { return p[n]; } };
struct Vec3 { Comp x, y, z; };
struct LbmView { Vec3 u; Comp rho; };
template<typename F> static void parallel_for(const ulong N, F f) { for(ulong n = 0ull; n < N; n++) f(n); } // FX/utilities.hpp:64-97 fans the same calls out to threads
struct View { float* p; float* data() { return p; } };
extern "C" void luwref_stats_accumulate(uint64_t N, uint64_t* count, const float* rho_, const float* u_, float* u_avg_, float* rho_avg_, float* m2_u_, float* m2_v_, float* m2_w_) {
	LbmView lbm; lbm.u.x.p = u_; lbm.u.y.p = u_ + N; lbm.u.z.p = u_ + 2u * N; lbm.rho.p = rho_;
	const ulong Ncells = N;
	ulong avg_count = *count;
	View avg_u{u_avg_}, avg_rho{rho_avg_}, M2_u{m2_u_}, M2_v{m2_v_}, M2_w{m2_w_};
#include "../_ref/stats_lambda.inc"
	*count = avg_count;
}
