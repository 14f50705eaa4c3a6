/*
 * ref_preamble.h -- TEST INFRASTRUCTURE.  The text OpenMM's CudaContext::createModule puts in
 * front of a kernel source before compiling it, reduced to what
 * /root/reference/platforms/cuda/src/kernels/constVLangevin.cu uses: the `real` / `mixed` typedef
 * families of the three precision models, make_real4 / make_mixed4 / make_mixed3, SQRT, RECIP and
 * USE_MIXED_PRECISION.  With it the reference's kernel source compiles UNMODIFIED, from where it
 * lies, both with nvcc for sm_100a (ref_cuda.cu) and with g++ for the host (ref_host.cpp, which
 * also emulates the CUDA execution model).  [OpenMM itself is absent from the image; the preamble
 * is determined by the identifiers the kernel source uses, OpenMM 7.x conventions recalled.]
 *
 * Select the model with -DCONSTV_REF_PREC=0|1|2 (single | mixed | double).  Each model is one
 * translation unit; the extern "C" kernel names get a per-model suffix so that all three link into
 * one library.
 */
#ifndef CONSTV_REF_PREAMBLE_H_
#define CONSTV_REF_PREAMBLE_H_

#ifndef CONSTV_REF_PREC
#error "define CONSTV_REF_PREC = 0 (single), 1 (mixed) or 2 (double)"
#endif

#define CONSTV_REF_CAT2(a, b) a##_##b
#define CONSTV_REF_CAT(a, b) CONSTV_REF_CAT2(a, b)
#if CONSTV_REF_PREC == 0
#define CONSTV_REF_SUFFIX single
#elif CONSTV_REF_PREC == 1
#define CONSTV_REF_SUFFIX mixed
#else
#define CONSTV_REF_SUFFIX double
#endif
#define CONSTV_REF_NAME(x) CONSTV_REF_CAT(x, CONSTV_REF_SUFFIX)

#define integrateConstVLangevinPart1 CONSTV_REF_NAME(integrateConstVLangevinPart1)
#define integrateConstVLangevinPart2 CONSTV_REF_NAME(integrateConstVLangevinPart2)
#define selectConstVLangevinStepSize CONSTV_REF_NAME(selectConstVLangevinStepSize)
#define updateImageParticlePositions CONSTV_REF_NAME(updateImageParticlePositions)
#define reorderInverseAtomOrderIndexes CONSTV_REF_NAME(reorderInverseAtomOrderIndexes)
#define integrateConstVDrudeLangevinPart1 CONSTV_REF_NAME(integrateConstVDrudeLangevinPart1)
#define integrateConstVDrudeLangevinPart2 CONSTV_REF_NAME(integrateConstVDrudeLangevinPart2)
#define applyHardWallConstraints CONSTV_REF_NAME(applyHardWallConstraints)

#ifndef __CUDACC__
/* ---- host build: the CUDA dialect the kernel source is written in ---------------------------- */
#include <algorithm>
#include <cmath>
struct int2 { int x, y; };
struct int3 { int x, y, z; };
struct int4 { int x, y, z, w; };
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct double3 { double x, y, z; };
struct double4 { double x, y, z, w; };
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int3 make_int3(int x, int y, int z) { return int3{x, y, z}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline double3 make_double3(double x, double y, double z) { return double3{x, y, z}; }
static inline double4 make_double4(double x, double y, double z, double w) { return double4{x, y, z, w}; }
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
#define __device__
struct ConstVRefDim3 { unsigned int x, y, z; };
/* one emulated CUDA thread per host thread at a time; grid and block shape are launch-wide */
static thread_local ConstVRefDim3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0};
static ConstVRefDim3 blockDim = {1, 1, 1}, gridDim = {1, 1, 1};
#define __global__
#define __shared__
#define __syncthreads() ((void) 0)
#define __CUDA_ARCH__ 1000 /* the kernel source selects its double-precision 1/dt branch with >= 130 */
using std::exp;
using std::fabs;
using std::min;
using std::sqrt;
/* `extern __shared__ mixed params[];` in the (never launched) step-size kernel: give it storage */
#define params CONSTV_REF_NAME(constv_ref_dynamic_shared)
#endif

#if CONSTV_REF_PREC == 2
typedef double real;
typedef double2 real2;
typedef double3 real3;
typedef double4 real4;
#define make_real3 make_double3
#define make_real4 make_double4
#define USE_DOUBLE_PRECISION 1
#else
typedef float real;
typedef float2 real2;
typedef float3 real3;
typedef float4 real4;
#define make_real3 make_float3
#define make_real4 make_float4
#endif

#if CONSTV_REF_PREC == 0
typedef float mixed;
typedef float2 mixed2;
typedef float3 mixed3;
typedef float4 mixed4;
#define make_mixed3 make_float3
#define make_mixed4 make_float4
#define SQRT sqrtf
#define RECIP 1.0f/
#else
typedef double mixed;
typedef double2 mixed2;
typedef double3 mixed3;
typedef double4 mixed4;
#define make_mixed3 make_double3
#define make_mixed4 make_double4
#define SQRT sqrt
#define RECIP 1.0/
#endif

#if CONSTV_REF_PREC == 1
#define USE_MIXED_PRECISION 1
#endif

#endif /* CONSTV_REF_PREAMBLE_H_ */
