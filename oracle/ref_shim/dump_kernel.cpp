// TEST INFRASTRUCTURE (oracle/): prints the reference's OpenCL-C program text.
// Linked against <reference>/core/cfd_core/FluidX3D/src/kernel.cpp, which defines get_opencl_c_code()
// through its own header (FX/kernel.hpp:6-17). Nothing of the reference is included here.
#include <cstdio>
#include <string>
std::string get_opencl_c_code();
int main() {
	const std::string s = get_opencl_c_code();
	fwrite(s.data(), 1, s.size(), stdout);
	return 0;
}
