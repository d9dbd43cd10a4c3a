// Minimal host stand-ins for the CUDA qualifiers and intrinsics used by the marked regions of
// lsml_b200/csrc/fmm.cu, so that g++ can compile the kernel's device functions unchanged.
// Test infrastructure only (tests/test_fmm_device_code_host.py).
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>

#define __device__
#define __host__
#define __forceinline__ inline

typedef unsigned char u8;
typedef long long i64;
typedef unsigned long long u64;

template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __double2float_rn(double d) { return (float)d; }      // round to nearest even
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
using std::max;
using std::min;

// ---- kernels without block-level cooperation run as plain functions on one host "thread" ----
#define __global__
#define __launch_bounds__(...)
struct HostDim3 { unsigned x, y, z; };
static HostDim3 blockIdx = {0, 0, 0}, threadIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};
