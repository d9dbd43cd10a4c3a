// Minimal host stand-ins for the CUDA qualifiers and intrinsics used by the marked regions of
// lsml_b200/csrc/fmm.cu, so that g++ can compile the kernel's device functions unchanged.
// Test infrastructure only (tests/test_fmm_device_code_host.py).
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

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

// ---- single-lane stand-ins for atomics, warp intrinsics and cooperative groups -----------------
static inline unsigned atomicOr(unsigned* p, unsigned v) { unsigned o = *p; *p = o | v; return o; }
static inline u64 atomicAdd(u64* p, u64 v) { u64 o = *p; *p = o + v; return o; }
static inline unsigned atomicMax(unsigned* p, unsigned v) { unsigned o = *p; if (v > o) *p = v; return o; }
static inline u64 atomicMax(u64* p, u64 v) { u64 o = *p; if (v > o) *p = v; return o; }
static inline unsigned __activemask() { return 1u; }
static inline unsigned __match_any_sync(unsigned, unsigned) { return 1u; }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
template <class T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
template <class T> static inline void __stcg(T* p, T v) { *p = v; }
struct uint2 { unsigned x, y; };
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r; r.x = x; r.y = y; return r; }
template <> inline uint2 __ldcg<uint2>(const uint2* p) { return *p; }
namespace cooperative_groups {
struct coalesced_group {
    unsigned size() const { return 1u; }
    unsigned thread_rank() const { return 0u; }
    template <class T> T shfl(T v, int) const { return v; }
};
static inline coalesced_group coalesced_threads() { return coalesced_group(); }
template <class T> static inline T inclusive_scan(const coalesced_group&, T v) { return v; }
}  // namespace cooperative_groups
namespace cg = cooperative_groups;
#define __align__(n) __attribute__((aligned(n)))
static inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
static inline unsigned __float_as_uint_(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
