// common.cuh -- shared helpers for the lsml_b200 kernels (sm_100a).
//
// Numerics contract: the reference path is float64 end to end and parity
// is judged on bit-exact band masks, so every kernel evaluates the reference's expressions in
// the reference's order.  The library is compiled with -fmad=false: no multiply-add
// contraction, IEEE add/mul/div/sqrt only, which makes device results bit-identical to the
// reference's gcc -O3 x86-64 build (which has no FMA either).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef unsigned char u8;
typedef long long i64;
typedef unsigned long long u64;

#define LSML_SMS 148  // B200: 2 dies x 74 SMs

namespace lsml {

// status / error reporting shared by the C ABI (capi.cu owns the storage)
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
void count_launch(int n = 1);

#define LSML_CUDA(call)                                             \
    do {                                                            \
        int _s = lsml::check_cuda((call), #call);                   \
        if (_s) return _s;                                          \
    } while (0)

#define LSML_LAUNCH_CHECK(name)                                     \
    do {                                                            \
        lsml::count_launch();                                       \
        int _s = lsml::check_cuda(cudaGetLastError(), name);        \
        if (_s) return _s;                                          \
    } while (0)

// Grid geometry, always padded to three axes (leading axes of extent 1 for ndim < 3).
// Axis a of an ndim grid is stored at index 3-ndim+a.
struct Geom {
    int ndim;
    int n[3];      // extents (padded)
    i64 st[3];     // element strides (padded)
    i64 voxels;    // per item
    i64 batch;
    double dx[3];  // spacing (padded with 1)
    double rdx[3]; // exact reciprocal when dx is a power of two, else 0 (=> divide)
};

inline bool is_pow2_double(double d) {
    if (!(d > 0)) return false;
    union { double f; uint64_t u; } v; v.f = d;
    return (v.u & 0x000fffffffffffffULL) == 0 && ((v.u >> 52) & 0x7ff) != 0 &&
           ((v.u >> 52) & 0x7ff) != 0x7ff;
}

inline int make_geom(Geom& g, int ndim, const int* shape, i64 batch, const double* dx) {
    if (ndim < 1 || ndim > 3) { set_error("ndim must be 1..3 (got %d)", ndim); return 1; }
    g.ndim = ndim; g.batch = batch;
    for (int a = 0; a < 3; ++a) { g.n[a] = 1; g.dx[a] = 1.0; g.rdx[a] = 1.0; }
    for (int a = 0; a < ndim; ++a) {
        if (shape[a] < 1) { set_error("shape[%d] = %d is not positive", a, shape[a]); return 1; }
        g.n[3 - ndim + a] = shape[a];
        double d = dx ? dx[a] : 1.0;
        g.dx[3 - ndim + a] = d;
        g.rdx[3 - ndim + a] = is_pow2_double(d) ? 1.0 / d : 0.0;
    }
    g.st[2] = 1; g.st[1] = g.n[2]; g.st[0] = (i64)g.n[1] * g.n[2];
    g.voxels = (i64)g.n[0] * g.n[1] * g.n[2];
    return 0;
}

// x / d, bit-exact: when d is a power of two the product with its exact reciprocal is the same
// correctly rounded value and costs one DMUL instead of a ~12-instruction division sequence.
template <typename T>
__device__ __forceinline__ T div_dx(T x, T d, T rd) {
    return rd != T(0) ? x * rd : x / d;
}

// helpers.c:57-58 (NaN in the first argument propagates)
template <typename T> __device__ __forceinline__ T ref_max(T a, T b) { return a < b ? b : a; }
template <typename T> __device__ __forceinline__ T ref_min(T a, T b) { return a > b ? b : a; }

__device__ __forceinline__ double dev_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float dev_sqrt(float x) { return sqrtf(x); }

// warp / block reductions (sum) for double and 64-bit integers
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the block; result valid in thread 0. `smem` must hold 32 elements of T.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* smem) {
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int lane = tid & 31, warp = tid >> 5;
    const int nwarps = (blockDim.x * blockDim.y * blockDim.z + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    T r = T(0);
    if (warp == 0) {
        r = lane < nwarps ? smem[lane] : T(0);
        r = warp_sum(r);
    }
    return r;
}

}  // namespace lsml
