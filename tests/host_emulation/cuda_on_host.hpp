// TEST INFRASTRUCTURE: just enough of the CUDA execution model to compile latticeurbanwind_b200/csrc/lbm_kernels.cuh with g++ and run its
// one-cell-per-thread kernels on host threads (tests/test_kernel_source_on_host.py). Purpose: check the LOGIC of kernel source that has not yet been
// observed on a B200 (indexing, slot parity, operation order) bit for bit against the oracle, in the container that has no GPU. It is not a CPU
// fallback: nothing in the package can reach it, and it says nothing about the compiled SASS (nvcc -fmad=false is matched with -ffp-contract=off).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
struct emu_uint3 { unsigned x, y, z; };
static thread_local emu_uint3 emu_blockIdx, emu_threadIdx, emu_blockDim, emu_gridDim;
#include <cuda_fp16.h> // host-callable __half conversions
#include <cuda_runtime.h>
#define blockIdx emu_blockIdx
#define threadIdx emu_threadIdx
#define blockDim emu_blockDim
#define gridDim emu_gridDim
#define __grid_constant__
#define __launch_bounds__(...)
#define __sincosf emu_sincosf
#define __syncthreads() ((void)0)
#undef __shared__
#define __shared__ static thread_local
#undef __constant__
#define __constant__ static const
template<typename T> static inline T __ldg(const T* p) { return *p; }
static inline float __uint_as_float(unsigned x) { float f; memcpy(&f, &x, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned x; memcpy(&x, &f, 4); return x; }
static inline float __fmul_rn(float a, float b) { return a*b; }
static inline void emu_sincosf(float x, float* s, float* c) { *s = sinf(x); *c = cosf(x); }
// packed / explicitly rounded FP32 intrinsics used by csrc/lbm_vec.cuh (each lane rounded like the scalar operation; g++ -ffp-contract=off keeps a*b+c unfused)
static inline float __fadd_rn(float a, float b) { return a+b; }
static inline float __fsub_rn(float a, float b) { return a-b; }
static inline float __fdiv_rn(float a, float b) { return a/b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float2 __fadd2_rn(float2 a, float2 b) { return make_float2(a.x+b.x, a.y+b.y); }
static inline float2 __fmul2_rn(float2 a, float2 b) { return make_float2(a.x*b.x, a.y*b.y); }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
