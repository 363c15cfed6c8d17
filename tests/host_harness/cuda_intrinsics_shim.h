/* TEST STUB -- the CUDA single-operation intrinsics the kernels use, as plain C for g++ -ffp-contract=off, so that
 * device functions that are pure arithmetic can be compiled and property-tested on the CPU (tests/test_raster_host.py).
 * Each is one IEEE binary32 operation, round to nearest even, exactly what the intrinsic is. */
#ifndef TB_TEST_CUDA_INTRINSICS_SHIM_H
#define TB_TEST_CUDA_INTRINSICS_SHIM_H
#include <cmath>
#include <cstdint>
#include <cstring>
#ifndef __forceinline__
#define __forceinline__ inline
#endif
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; std::memcpy(&i, &f, 4); return i; }
#endif
