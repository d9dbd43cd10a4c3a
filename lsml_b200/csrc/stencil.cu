// stencil.cu -- dense masked stencils on 1/2/3-D grids.
//
//   gmag_os_kernel            Osher-Sethian upwind |Du|      masked_gradient.c:152-315
//   gradient_centered_kernel  centered differences + |Du|    masked_gradient.c:13-150
//
// HBM-bound: algorithmic bytes per voxel (f64) = 8 (A) + 1 (mask) + 8 (nu) + 8 (gmag) = 25 for
// gmag_os and 8 + 1 + 8*ndim + 8 for gradient_centered.
//
// Layout / mapping: the grid is padded to (n0, n1, n2) with n2 the contiguous axis.  A CTA is
// BX x BY threads: BX consecutive voxels along n2 (coalesced 8-byte accesses, 256 B per warp
// row) times BY consecutive rows (n1 direction), so that the +-n1 neighbours of most voxels
// are loaded by the same CTA and hit L1; +-n0 neighbours are one plane away and hit L2
// (a 256^2 f64 plane is 512 KB; L2 is 126 MB).  Each voxel of A is therefore fetched from
// HBM once.  Masked-out voxels issue no loads at all (band masks touch ~2% of the grid).
#include <cstdlib>
#include <cuda.h>
#include "common.cuh"

namespace lsml {

struct StencilParams {
    int n0, n1, n2;
    i64 st0, st1;          // st2 == 1
    i64 rows;              // batch * n0 * n1
    double dx0, dx1, dx2;
    double r0, r1, r2;     // exact reciprocals or 0
};

// one-sided / two-sided differences along one axis (masked_gradient.c:165-176).
// Branch-free loads: the neighbour indices are clamped to the grid so that BOTH loads of every
// axis can be issued before anything is consumed (9 independent loads in flight per voxel
// instead of 4 dependent load phases -- the kernel was latency-bound, not bandwidth-bound);
// the border rule (b := f at index 0, f := b at index n-1) is applied afterwards by selects.
template <typename T>
__device__ __forceinline__ void load_pair(const T* __restrict__ A, int st, int c, int n,
                                          T& plus, T& minus) {
    plus = __ldg(A + (c + 1 < n ? st : 0));
    minus = __ldg(A + (c > 0 ? -st : 0));
}

// spacing: EXACT = every dx is a power of two => multiply by the exact reciprocal (bit-identical
// to the division, one DMUL); otherwise the IEEE division.  A compile-time switch keeps the
// ~25-instruction division sequence out of the common (dx = 1) kernel altogether.
template <typename T, bool EXACT>
__device__ __forceinline__ T scale_dx(T x, T d, T rd) {
    return EXACT ? x * rd : x / d;
}
template <typename T>
__device__ __forceinline__ void fwd_bwd(T plus, T minus, int c, int n, T center, T& f, T& b) {
    f = plus - center;
    b = center - minus;
    if (c == 0) b = f;                 // extent-1 axes: plus == minus == center => f = b = 0
    else if (c == n - 1) f = b;        // (the reference reads out of bounds there)
}

template <typename T>
__device__ __forceinline__ T os_terms(T b, T f, bool neg, T acc, bool first) {
    T tb, tf;
    if (neg) { tb = ref_max(b, T(0)); tf = ref_min(f, T(0)); }
    else     { tb = ref_min(b, T(0)); tf = ref_max(f, T(0)); }
    tb = tb * tb; tf = tf * tf;
    return first ? (tb + tf) : ((acc + tb) + tf);
}

template <typename T, int NDIM, bool EXACT>
__global__ void __launch_bounds__(256)
gmag_os_kernel(const T* __restrict__ A, const u8* __restrict__ mask, const T* __restrict__ nu,
               T* __restrict__ out, StencilParams p) {
    // blockIdx.x = item*n0 + i (one plane per CTA row), blockIdx.y = j tile, blockIdx.z = k tile:
    // no per-thread integer division (64-bit div/mod cost ~200 instructions per voxel before)
    const int k = blockIdx.z * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (k >= p.n2 || j >= p.n1) return;
    const int i = (int)(blockIdx.x % (unsigned)p.n0);
    const i64 l = ((i64)blockIdx.x * p.n1 + j) * p.n2 + k;
    if (mask != nullptr && !mask[l]) return;
    const T* __restrict__ Al = A + l;     // in-plane / plane strides fit 32 bits
    const T c = __ldg(Al);
    const T nuv = __ldg(nu + l);
    T p0 = c, m0 = c, p1 = c, m1 = c, p2, m2;
    if (NDIM >= 3) load_pair(Al, (int)p.st0, i, p.n0, p0, m0);
    if (NDIM >= 2) load_pair(Al, (int)p.st1, j, p.n1, p1, m1);
    load_pair(Al, 1, k, p.n2, p2, m2);
    const bool neg = nuv < T(0);
    T acc = T(0), f, b;
    bool first = true;
    if (NDIM >= 3) {
        fwd_bwd(p0, m0, i, p.n0, c, f, b);
        f = scale_dx<T, EXACT>(f, (T)p.dx0, (T)p.r0); b = scale_dx<T, EXACT>(b, (T)p.dx0, (T)p.r0);
        acc = os_terms(b, f, neg, acc, first); first = false;
    }
    if (NDIM >= 2) {
        fwd_bwd(p1, m1, j, p.n1, c, f, b);
        f = scale_dx<T, EXACT>(f, (T)p.dx1, (T)p.r1); b = scale_dx<T, EXACT>(b, (T)p.dx1, (T)p.r1);
        acc = os_terms(b, f, neg, acc, first); first = false;
    }
    fwd_bwd(p2, m2, k, p.n2, c, f, b);
    f = scale_dx<T, EXACT>(f, (T)p.dx2, (T)p.r2); b = scale_dx<T, EXACT>(b, (T)p.dx2, (T)p.r2);
    acc = os_terms(b, f, neg, acc, first);
    out[l] = dev_sqrt(acc);
}

// centered difference along one axis (masked_gradient.c:24-60), from the clamped pair
template <typename T>
__device__ __forceinline__ T centered(T plus, T minus, int c, int n, T center) {
    if (n == 1) return T(0);
    if (c == 0) return plus - center;
    if (c == n - 1) return center - minus;
    return T(0.5) * (plus - minus);
}

template <typename T, int NDIM, bool EXACT>
__global__ void __launch_bounds__(256)
gradient_centered_kernel(const T* __restrict__ A, const u8* __restrict__ mask,
                         T* __restrict__ grads, T* __restrict__ gmag, StencilParams p,
                         i64 plane, int normalize) {
    // blockIdx.x = item*n0 + i (one plane per CTA row), blockIdx.y = j tile, blockIdx.z = k tile:
    // no per-thread integer division (64-bit div/mod cost ~200 instructions per voxel before)
    const int k = blockIdx.z * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (k >= p.n2 || j >= p.n1) return;
    const int i = (int)(blockIdx.x % (unsigned)p.n0);
    const i64 l = ((i64)blockIdx.x * p.n1 + j) * p.n2 + k;
    if (mask != nullptr && !mask[l]) return;
    const T c = __ldg(A + l);
    const T* __restrict__ Al = A + l;
    T p0 = c, m0 = c, p1 = c, m1 = c, p2, m2;
    if (NDIM >= 3) load_pair(Al, (int)p.st0, i, p.n0, p0, m0);
    if (NDIM >= 2) load_pair(Al, (int)p.st1, j, p.n1, p1, m1);
    load_pair(Al, 1, k, p.n2, p2, m2);
    T d0 = T(0), d1 = T(0), d2;
    if (NDIM >= 3) d0 = scale_dx<T, EXACT>(centered(p0, m0, i, p.n0, c), (T)p.dx0, (T)p.r0);
    if (NDIM >= 2) d1 = scale_dx<T, EXACT>(centered(p1, m1, j, p.n1, c), (T)p.dx1, (T)p.r1);
    d2 = scale_dx<T, EXACT>(centered(p2, m2, k, p.n2, c), (T)p.dx2, (T)p.r2);
    T g;
    if (NDIM == 1) g = d2 > T(0) ? d2 : -d2;                       // masked_gradient.c:144
    else if (NDIM == 2) g = dev_sqrt(d1 * d1 + d2 * d2);           // :113
    else g = dev_sqrt((d0 * d0 + d1 * d1) + d2 * d2);              // :65
    gmag[l] = g;
    if (normalize == 1 && g > T(0)) { d0 /= g; d1 /= g; d2 /= g; }
    if (NDIM == 3) { grads[l] = d0; grads[plane + l] = d1; grads[2 * plane + l] = d2; }
    else if (NDIM == 2) { grads[l] = d1; grads[plane + l] = d2; }
    else grads[l] = d2;
}

// ---------------------------------------------------------------------------------------------
// 3-D column-marching kernels (the 256^3 / 1024^3 dense case).
//
// The one-voxel-per-thread kernels above are ISSUE-bound on B200 (ncu: issue slots 69% busy,
// DRAM 23%, ~300 instructions per voxel, most of them 64-bit address arithmetic and border
// selects).  Here a thread owns one (j, k) column of a chunk of planes and marches along axis 0:
// A[i-1], A[i], A[i+1] live in registers (one new global load of A per voxel for that axis), the
// four in-plane neighbour offsets and border flags are computed once per thread, and all
// addresses are 32-bit offsets from per-CTA base pointers.  ~85 instructions per voxel, which
// puts the kernel on the HBM roofline (25 B/voxel) instead of the issue limit.
// UNIT: every dx == 1 (x/1 == x exactly, the scaling multiply is dropped).
// ---------------------------------------------------------------------------------------------
constexpr int MARCH_BX = 64, MARCH_BY = 8, MARCH_CHUNK = 16;

// sign flip by XOR on the high word: (neg ? -x : x) in one LOP3 instead of a two-FSEL select
__device__ __forceinline__ double flip_sign(double x, unsigned sgn_hi) {
    return __hiloint2double(__double2hiint(x) ^ (int)sgn_hi, __double2loint(x));
}
__device__ __forceinline__ float flip_sign(float x, unsigned sgn_hi) {
    return __int_as_float(__float_as_int(x) ^ (int)sgn_hi);
}

// One axis of the Osher-Sethian sum.  With s = (nu < 0 ? -1 : +1) the reference's
//   nu <  0: max(b,0)^2 + min(f,0)^2      nu >= 0: min(b,0)^2 + max(f,0)^2
// is min(s*b, 0)^2 + max(s*f, 0)^2 for both signs (negation is exact, the squares are equal
// and NaNs propagate through the same comparisons as helpers.c:57-58).
template <typename T>
__device__ __forceinline__ T os_axis(T b, T f, unsigned sgn_hi, T acc, bool first) {
    const T sb = flip_sign(b, sgn_hi), sf = flip_sign(f, sgn_hi);
    const T tb = sb > T(0) ? T(0) : sb;
    const T tf = sf < T(0) ? T(0) : sf;
    return first ? (tb * tb + tf * tf) : ((acc + tb * tb) + tf * tf);
}

// sqrt(0) == 0 without entering the library's slow path (local extrema of u give exactly 0, and
// the library handles a zero argument in a called subroutine that the whole warp then walks
// through): the root is taken of a harmless stand-in and discarded.
__device__ __forceinline__ void opaque(double& x) { asm volatile("" : "+d"(x)); }
__device__ __forceinline__ void opaque(float& x) { asm volatile("" : "+f"(x)); }
template <typename T>
__device__ __forceinline__ T sqrt_nonneg(T x) {
    const bool pos = x > T(0);
    T arg = pos ? x : T(1);
    opaque(arg);                 // or the compiler hoists sqrt(x) above the select again
    const T r = dev_sqrt(arg);
    return pos ? r : x;
}

template <typename T, bool EXACT, bool UNIT>
__global__ void __launch_bounds__(MARCH_BX * MARCH_BY)
gmag_os_march3d_kernel(const T* __restrict__ A, const u8* __restrict__ mask,
                       const T* __restrict__ nu, T* __restrict__ out, StencilParams p) {
    const int k = blockIdx.z * MARCH_BX + threadIdx.x;
    const int j = blockIdx.y * MARCH_BY + threadIdx.y;
    if (k >= p.n2 || j >= p.n1) return;
    const unsigned chunks = (unsigned)((p.n0 + MARCH_CHUNK - 1) / MARCH_CHUNK);
    const unsigned item = blockIdx.x / chunks, ci = blockIdx.x % chunks;
    const int i0 = (int)ci * MARCH_CHUNK;
    const int i1 = min(p.n0, i0 + MARCH_CHUNK);
    const i64 st0 = p.st0;
    const i64 first = ((i64)item * p.n0 + i0) * st0 + (i64)j * p.n2 + k;
    // running pointers: one 64-bit multiply-add each per plane instead of per-access indexing
    const T* __restrict__ pc = A + first;
    const T* __restrict__ pjp = pc + ((j + 1 < p.n1) ? p.n2 : 0);
    const T* __restrict__ pjm = pc - ((j > 0) ? p.n2 : 0);
    const T* __restrict__ pkp = pc + ((k + 1 < p.n2) ? 1 : 0);
    const T* __restrict__ pkm = pc - ((k > 0) ? 1 : 0);
    const T* __restrict__ pnu = nu + first;
    const u8* __restrict__ pm = mask ? mask + first : nullptr;
    T* __restrict__ po = out + first;
    const bool jlo = j == 0, jhi = j == p.n1 - 1, klo = k == 0, khi = k == p.n2 - 1;
    const bool border_jk = jlo | jhi | klo | khi;
    T cur = __ldg(pc);
    T prv = (i0 > 0) ? __ldg(pc - st0) : cur;
    for (int i = i0; i < i1; ++i) {
        const T nxt = (i + 1 < p.n0) ? __ldg(pc + st0) : cur;
        if (pm == nullptr || *pm) {
            const T jp = __ldg(pjp), jm = __ldg(pjm), kp = __ldg(pkp), km = __ldg(pkm);
            const unsigned sgn = (__ldg(pnu) < T(0)) ? 0x80000000u : 0u;
            T f0 = nxt - cur, b0 = cur - prv;
            if (i == 0) b0 = f0; else if (i == p.n0 - 1) f0 = b0;     // uniform
            T f1 = jp - cur, b1 = cur - jm, f2 = kp - cur, b2 = cur - km;
            if (border_jk) {                                           // rare lanes only
                if (jlo) b1 = f1; else if (jhi) f1 = b1;
                if (klo) b2 = f2; else if (khi) f2 = b2;
            }
            if (!UNIT) {
                f0 = scale_dx<T, EXACT>(f0, (T)p.dx0, (T)p.r0); b0 = scale_dx<T, EXACT>(b0, (T)p.dx0, (T)p.r0);
                f1 = scale_dx<T, EXACT>(f1, (T)p.dx1, (T)p.r1); b1 = scale_dx<T, EXACT>(b1, (T)p.dx1, (T)p.r1);
                f2 = scale_dx<T, EXACT>(f2, (T)p.dx2, (T)p.r2); b2 = scale_dx<T, EXACT>(b2, (T)p.dx2, (T)p.r2);
            }
            T acc = os_axis(b0, f0, sgn, T(0), true);
            acc = os_axis(b1, f1, sgn, acc, false);
            acc = os_axis(b2, f2, sgn, acc, false);
            *po = sqrt_nonneg(acc);
        }
        prv = cur; cur = nxt;
        pc += st0; pjp += st0; pjm += st0; pkp += st0; pkm += st0; pnu += st0; po += st0;
        if (pm) pm += st0;
    }
}

template <typename T, bool EXACT, bool UNIT>
__global__ void __launch_bounds__(MARCH_BX * MARCH_BY)
gradient_centered_march3d_kernel(const T* __restrict__ A, const u8* __restrict__ mask,
                                 T* __restrict__ grads, T* __restrict__ gmag, StencilParams p,
                                 i64 plane, int normalize) {
    const int k = blockIdx.z * MARCH_BX + threadIdx.x;
    const int j = blockIdx.y * MARCH_BY + threadIdx.y;
    if (k >= p.n2 || j >= p.n1) return;
    const unsigned chunks = (unsigned)((p.n0 + MARCH_CHUNK - 1) / MARCH_CHUNK);
    const unsigned item = blockIdx.x / chunks, ci = blockIdx.x % chunks;
    const int i0 = (int)ci * MARCH_CHUNK;
    const int i1 = min(p.n0, i0 + MARCH_CHUNK);
    const int st0 = (int)p.st0;
    const i64 base = ((i64)item * p.n0 + i0) * p.st0;
    const T* __restrict__ Ab = A + base;
    const u8* __restrict__ mb = mask ? mask + base : nullptr;
    T* __restrict__ g0 = grads + base;
    T* __restrict__ g1 = grads + plane + base;
    T* __restrict__ g2 = grads + 2 * plane + base;
    T* __restrict__ gm = gmag + base;
    const int ojp = (j + 1 < p.n1) ? p.n2 : 0, ojm = (j > 0) ? -p.n2 : 0;
    const int okp = (k + 1 < p.n2) ? 1 : 0, okm = (k > 0) ? -1 : 0;
    const bool jint = j > 0 && j < p.n1 - 1, kint = k > 0 && k < p.n2 - 1;
    int off = j * p.n2 + k;
    T cur = __ldg(Ab + off);
    T prv = (i0 > 0) ? __ldg(Ab + off - st0) : cur;
#pragma unroll 2
    for (int i = i0; i < i1; ++i, off += st0) {
        const T nxt = (i + 1 < p.n0) ? __ldg(Ab + off + st0) : cur;
        if (mb == nullptr || mb[off]) {
            const T jp = __ldg(Ab + off + ojp), jm = __ldg(Ab + off + ojm);
            const T kp = __ldg(Ab + off + okp), km = __ldg(Ab + off + okm);
            // at a border one of the pair equals `cur`, so (plus - minus) is the one-sided
            // difference there (masked_gradient.c:24-60); extent-1 axes give 0
            T d0 = nxt - prv, d1 = jp - jm, d2 = kp - km;
            if (i > 0 && i < p.n0 - 1) d0 = T(0.5) * d0;
            if (jint) d1 = T(0.5) * d1;
            if (kint) d2 = T(0.5) * d2;
            if (!UNIT) {
                d0 = scale_dx<T, EXACT>(d0, (T)p.dx0, (T)p.r0);
                d1 = scale_dx<T, EXACT>(d1, (T)p.dx1, (T)p.r1);
                d2 = scale_dx<T, EXACT>(d2, (T)p.dx2, (T)p.r2);
            }
            const T g = dev_sqrt((d0 * d0 + d1 * d1) + d2 * d2);
            gm[off] = g;
            if (normalize == 1 && g > T(0)) { d0 /= g; d1 /= g; d2 /= g; }
            g0[off] = d0; g1[off] = d1; g2[off] = d2;
        }
        prv = cur; cur = nxt;
    }
}

// ---------------------------------------------------------------------------------------------
// Vectorised column march (the 256^3 / 1024^3 float64 and float32 case with an even last axis).
//
// ncu on the scalar march above: issue slots 68 % busy, DRAM 39 % -- ~200 instructions per voxel,
// most of them compare/select pairs of the upwind switch, 64-bit pointer bumps and one memory
// instruction per operand.  Here a thread owns VEC consecutive voxels along the contiguous axis
// (16-byte loads / stores: half or a quarter of the memory instructions and pointer updates per
// voxel; the k-neighbours of the inner voxels are already in registers), and the upwind terms
// use the identities
//     min(x, 0)^2 = ((x - |x|) / 2)^2        max(x, 0)^2 = ((x + |x|) / 2)^2
// with the factor 1/4 pulled out of the sum and 1/2 out of the square root:
//     |Du| = 0.5 * sqrt( sum (x -+ |x|)^2 ).
// x -+ |x| is exact (0 or 2x), scaling by powers of two commutes with every rounding of the
// reference's left-to-right sum and with sqrt, so the result is BIT-IDENTICAL to the reference
// expression whenever no intermediate is subnormal, infinite or NaN.  A cheap range test on the
// sum detects everything else (sum not finite, or 0 < sum < 2^-900) and re-evaluates that voxel
// with the reference's own expressions (os_terms), so the kernel is exact for all inputs.
// ---------------------------------------------------------------------------------------------
template <typename T> struct VecOf;
template <> struct VecOf<double> { typedef double2 type; static constexpr int N = 2; };
template <> struct VecOf<float> { typedef float4 type; static constexpr int N = 4; };

constexpr int MV_BX = 32, MV_BY = 8, MV_CHUNK = 16;
#ifndef MV_MIN_BLOCKS
#define MV_MIN_BLOCKS 3
#endif

__device__ __forceinline__ void unpack(const double2& v, double (&a)[2]) { a[0] = v.x; a[1] = v.y; }
__device__ __forceinline__ void unpack(const float4& v, float (&a)[4]) { a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w; }
__device__ __forceinline__ double2 pack(const double (&a)[2]) { return make_double2(a[0], a[1]); }
__device__ __forceinline__ float4 pack(const float (&a)[4]) { return make_float4(a[0], a[1], a[2], a[3]); }

// 4 * min(x,0)^2 and 4 * max(x,0)^2
template <typename T> __device__ __forceinline__ T neg_part_sq4(T x) { const T t = x - fabs(x); return t * t; }
template <typename T> __device__ __forceinline__ T pos_part_sq4(T x) { const T t = x + fabs(x); return t * t; }

// true when 4*sum is safely inside the range where the scaled evaluation is exact
__device__ __forceinline__ bool scaled_sum_ok(double s4) {
    const unsigned hi = (unsigned)__double2hiint(s4);       // s4 >= 0 or NaN
    return s4 == 0.0 || (hi - 0x07b00000u) < (0x7ff00000u - 0x07b00000u);   // 2^-900 <= s4 < inf
}
__device__ __forceinline__ bool scaled_sum_ok(float s4) {
    const unsigned b = __float_as_uint(s4);
    return s4 == 0.0f || (b - 0x0d800000u) < (0x7f800000u - 0x0d800000u);   // 2^-100 <= s4 < inf
}

// the reference's own expression for one voxel (slow path of the range test, and borders)
template <typename T>
__device__ __noinline__ T gmag_os_exact(T b0, T f0, T b1, T f1, T b2, T f2, bool neg) {
    T acc = os_terms(b0, f0, neg, T(0), true);
    acc = os_terms(b1, f1, neg, acc, false);
    acc = os_terms(b2, f2, neg, acc, false);
    return dev_sqrt(acc);
}

template <typename T, bool EXACT, bool UNIT>
__global__ void __launch_bounds__(MV_BX * MV_BY, MV_MIN_BLOCKS)
gmag_os_marchv_kernel(const T* __restrict__ A, const u8* __restrict__ mask,
                      const T* __restrict__ nu, T* __restrict__ out, StencilParams p) {
    typedef typename VecOf<T>::type V;
    constexpr int N = VecOf<T>::N;
    const int k0 = (blockIdx.z * MV_BX + threadIdx.x) * N;
    const int j = blockIdx.y * MV_BY + threadIdx.y;
    if (k0 >= p.n2 || j >= p.n1) return;
    const unsigned chunks = (unsigned)((p.n0 + MV_CHUNK - 1) / MV_CHUNK);
    const unsigned item = blockIdx.x / chunks, ci = blockIdx.x % chunks;
    const int i0 = (int)ci * MV_CHUNK;
    const int i1 = min(p.n0, i0 + MV_CHUNK);
    const i64 st0 = p.st0;
    const i64 first = ((i64)item * p.n0 + i0) * st0 + (i64)j * p.n2 + k0;
    const T* __restrict__ pc = A + first;
    const T* __restrict__ pnu = nu + first;
    const u8* __restrict__ pm = mask ? mask + first : nullptr;
    T* __restrict__ po = out + first;
    const bool jlo = j == 0, jhi = j == p.n1 - 1, klo = k0 == 0, khi = k0 + N == p.n2;
    const int ojp = jhi ? 0 : p.n2, ojm = jlo ? 0 : -p.n2;       // clamped: the border rule below
    const int okp = khi ? N - 1 : N, okm = klo ? 0 : -1;          // overrides what they fetch
    const bool border_jk = jlo | jhi | klo | khi;
    auto ldv = [](const T* q) { return __ldg(reinterpret_cast<const V*>(q)); };
    // raw mask bytes of the thread's N voxels (all-ones without a mask); they are only LOOKED at
    // an iteration after they were requested, so the load never stalls the pipeline
    auto ldmask = [](const u8* q) -> unsigned {
        if (q == nullptr) return 0x01010101u >> (8 * (4 - N));
        if (N == 2) return __ldg(reinterpret_cast<const unsigned short*>(q));
        return __ldg(reinterpret_cast<const unsigned*>(q));
    };
    auto mask_bits = [](unsigned raw) -> unsigned {
        unsigned bits = 0u;
#pragma unroll
        for (int v = 0; v < N; ++v) bits |= ((raw >> (8 * v)) & 0xffu) ? (1u << v) : 0u;
        return bits;
    };
    // Software pipeline (the kernel is bound by bytes in flight, not by instructions): the HBM
    // operands of a plane are requested before they are needed -- A two planes ahead, the mask
    // two planes ahead, nu one plane ahead (only for planes with a masked voxel, and never
    // waiting for the mask).  The in-plane neighbours are L1/L2 hits and are fetched in place.
    V cur = ldv(pc);
    V prv = i0 > 0 ? ldv(pc - st0) : cur;
    V nxt = i0 + 1 < p.n0 ? ldv(pc + st0) : cur;
    unsigned raw_cur = ldmask(pm);
    unsigned raw_nxt = i0 + 1 < i1 ? ldmask(pm ? pm + st0 : nullptr) : 0u;
    V nu_cur = cur;
    if (raw_cur) nu_cur = ldv(pnu);
    for (int i = i0; i < i1; ++i) {
        V nxt2 = nxt;
        if (i + 2 < p.n0) nxt2 = ldv(pc + 2 * st0);
        const unsigned raw_nxt2 = i + 2 < i1 ? ldmask(pm ? pm + 2 * st0 : nullptr) : 0u;
        V nu_nxt = nu_cur;
        if (raw_nxt) nu_nxt = ldv(pnu + st0);
        const unsigned m_cur = mask_bits(raw_cur);
        if (m_cur) {
            T c[N], pv[N], nx[N], jp[N], jm[N], nuv[N], res[N];
            unpack(ldv(pc + ojp), jp);
            unpack(ldv(pc + ojm), jm);
            const T right = __ldg(pc + okp), left = __ldg(pc + okm);
            unpack(cur, c); unpack(prv, pv); unpack(nxt, nx); unpack(nu_cur, nuv);
            const bool ilo = i == 0, ihi = i == p.n0 - 1;
#pragma unroll
            for (int v = 0; v < N; ++v) {
                const T cc = c[v];
                T f0 = nx[v] - cc, b0 = cc - pv[v];
                T f1 = jp[v] - cc, b1 = cc - jm[v];
                T f2 = (v + 1 < N ? c[v + 1 < N ? v + 1 : v] : right) - cc;
                T b2 = cc - (v > 0 ? c[v > 0 ? v - 1 : 0] : left);
                if (ilo) b0 = f0; else if (ihi) f0 = b0;                // uniform
                if (border_jk) {                                        // rare threads only
                    if (jlo) b1 = f1; else if (jhi) f1 = b1;
                    if (v == 0 && klo) b2 = f2;
                    if (v == N - 1 && khi) f2 = b2;
                }
                if (!UNIT) {
                    f0 = scale_dx<T, EXACT>(f0, (T)p.dx0, (T)p.r0); b0 = scale_dx<T, EXACT>(b0, (T)p.dx0, (T)p.r0);
                    f1 = scale_dx<T, EXACT>(f1, (T)p.dx1, (T)p.r1); b1 = scale_dx<T, EXACT>(b1, (T)p.dx1, (T)p.r1);
                    f2 = scale_dx<T, EXACT>(f2, (T)p.dx2, (T)p.r2); b2 = scale_dx<T, EXACT>(b2, (T)p.dx2, (T)p.r2);
                }
                const bool neg = nuv[v] < T(0);
                const unsigned sgn = neg ? 0x80000000u : 0u;
                // s = -1 where nu < 0: min(s*b, 0)^2 + max(s*f, 0)^2 per axis (see os_axis)
                T s4 = neg_part_sq4(flip_sign(b0, sgn)) + pos_part_sq4(flip_sign(f0, sgn));
                s4 = (s4 + neg_part_sq4(flip_sign(b1, sgn))) + pos_part_sq4(flip_sign(f1, sgn));
                s4 = (s4 + neg_part_sq4(flip_sign(b2, sgn))) + pos_part_sq4(flip_sign(f2, sgn));
                T r = T(0.5) * sqrt_nonneg(s4);
                if (!scaled_sum_ok(s4)) r = gmag_os_exact<T>(b0, f0, b1, f1, b2, f2, neg);
                res[v] = r;
            }
            if (m_cur == (1u << N) - 1u) {
                *reinterpret_cast<V*>(po) = pack(res);
            } else {
#pragma unroll
                for (int v = 0; v < N; ++v)
                    if ((m_cur >> v) & 1u) po[v] = res[v];
            }
        }
        prv = cur; cur = nxt; nxt = nxt2;
        raw_cur = raw_nxt; raw_nxt = raw_nxt2; nu_cur = nu_nxt;
        pc += st0; pnu += st0; po += st0;
        if (pm) pm += st0;
    }
}

// ---------------------------------------------------------------------------------------------
// TMA plane pipeline (the 256^3 / 1024^3 case, float64 and float32; north_star: "shared-memory or
// TMA tile staging, halos").
//
// ncu on the vectorised march above: 311 instructions per two-voxel iteration, issue slots 49 %
// busy at 24 warps/SM, DRAM 39 % -- more than half of the instructions are 64-bit address
// arithmetic, pointer bumps and the seven global loads per iteration.  Here NO consumer thread
// computes a global load address: a CTA owns a (TJ x TK) column tile of a chunk of planes and
// marches along axis 0 while one producer warp streams, per plane,
//     the A tile WITH its halo   (TJ+2) rows x (512 + 32) bytes   cp.async.bulk.tensor.4d
//     the nu tile                 TJ rows x 512 bytes
//     the mask tile               TJ rows x TK bytes
// into a ring of shared-memory stages through three tensor maps over (k, j, i, item); an
// mbarrier per stage carries the byte count (full) and the consumers' release (empty), so the
// eight consumer warps never meet at a block barrier.  Out-of-range halo cells are zero-filled
// by the TMA unit and then overridden by the reference's border rule (b := f at index 0,
// f := b at index n-1; masked_gradient.c:165-176) exactly as in the kernels above.  Every
// stencil neighbour is a shared-memory load at a CONSTANT offset from one base register
// (16-byte LDS for the pairs along k: the halo is two cells wide on the left to keep them
// aligned).  A[i-1] and A[i] of the column stay in registers.  The arithmetic is the vectorised
// kernel's (bit-exact, same range test and fallback).
// ---------------------------------------------------------------------------------------------
// Tile geometry in BYTES is the same for both element types: a thread owns 16 bytes along k (two
// float64 or four float32 voxels), a warp one row of 512 bytes, a CTA eight rows.
constexpr int TM_TJ = 8;                        // rows per tile (one consumer warp each)
constexpr int TM_ROW_BYTES = 512;               // 64 float64 / 128 float32 voxels
constexpr int TM_AROW_BYTES = TM_ROW_BYTES + 32;  // A row in smem: 16 bytes of halo on either side
#ifndef TM_STAGES_N
#define TM_STAGES_N 4
#endif
#ifndef TM_MIN_CTAS
#define TM_MIN_CTAS 3
#endif
constexpr int TM_STAGES = TM_STAGES_N;          // nu / mask ring
constexpr int TM_A_RING = TM_STAGES + 2;        // A planes i-1, i and the STAGES planes in flight
constexpr int TM_A_BYTES = (TM_TJ + 2) * TM_AROW_BYTES;    // 5440 bytes land per A tile ...
constexpr int TM_A_SLOT = (TM_A_BYTES + 127) / 128 * 128;  // ... in 128-byte aligned ring slots
constexpr int TM_NU_BYTES = TM_TJ * TM_ROW_BYTES;          // 4096
constexpr int TM_M_MAX = TM_TJ * (TM_ROW_BYTES / 4);       // mask tile: 512 (f64) / 1024 (f32) bytes
constexpr int TM_CONSUMERS = TM_TJ * 32;                   // 256 threads
constexpr int TM_THREADS = TM_CONSUMERS + 32;              // + the producer warp
constexpr size_t TM_SMEM = (size_t)TM_A_RING * TM_A_SLOT + (size_t)TM_STAGES * (TM_NU_BYTES + TM_M_MAX) +
                           2 * (TM_STAGES + 1) * 8 + 128;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LSML_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LSML_DONE;\n"
        "bra LSML_WAIT;\n"
        "LSML_DONE:\n"
        "}\n" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(unsigned dst, const void* map, unsigned bar, int c0,
                                            int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5, %6}], [%2];"
        :: "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

struct TmaMaps {           // CUtensorMap is 128 bytes, 64-byte aligned
    alignas(64) unsigned char a[128];
    alignas(64) unsigned char nu[128];
    alignas(64) unsigned char m[128];
};

// one thread's N voxels: centre vector `cur` with its six neighbour sets
template <typename T>
struct TmaStencil {
    static constexpr int N = VecOf<T>::N;
    T prv[N], cur[N], nxt[N], jp[N], jm[N];
    T left, right;
};

// Osher-Sethian |Du| of the N voxels (vectorised kernel's arithmetic).  BORDER: apply the i / j
// border rule; the k border rule (only the first / last voxel of a row can need it) always is.
template <typename T, bool EXACT, bool UNIT, bool BORDER>
__device__ __forceinline__ void tma_upwind(const TmaStencil<T>& q, const T (&nuv)[VecOf<T>::N],
                                           const StencilParams& p, bool ilo, bool ihi, bool jlo,
                                           bool jhi, bool klo, bool khi, T (&res)[VecOf<T>::N]) {
    constexpr int N = VecOf<T>::N;
#pragma unroll
    for (int v = 0; v < N; ++v) {
        const T cc = q.cur[v];
        T f0 = q.nxt[v] - cc, b0 = cc - q.prv[v];
        T f1 = q.jp[v] - cc, b1 = cc - q.jm[v];
        T f2 = (v + 1 < N ? q.cur[v + 1 < N ? v + 1 : v] : q.right) - cc;
        T b2 = cc - (v > 0 ? q.cur[v > 0 ? v - 1 : 0] : q.left);
        if (BORDER) {
            if (ilo) b0 = f0; else if (ihi) f0 = b0;
            if (jlo) b1 = f1; else if (jhi) f1 = b1;
        }
        if (v == 0 && klo) b2 = f2;
        if (v == N - 1 && khi) f2 = b2;
        if (!UNIT) {
            f0 = scale_dx<T, EXACT>(f0, (T)p.dx0, (T)p.r0); b0 = scale_dx<T, EXACT>(b0, (T)p.dx0, (T)p.r0);
            f1 = scale_dx<T, EXACT>(f1, (T)p.dx1, (T)p.r1); b1 = scale_dx<T, EXACT>(b1, (T)p.dx1, (T)p.r1);
            f2 = scale_dx<T, EXACT>(f2, (T)p.dx2, (T)p.r2); b2 = scale_dx<T, EXACT>(b2, (T)p.dx2, (T)p.r2);
        }
        const bool neg = nuv[v] < T(0);
        const unsigned sgn = neg ? 0x80000000u : 0u;
        // s = -1 where nu < 0: min(s*b, 0)^2 + max(s*f, 0)^2 per axis (see os_axis)
        T s4 = neg_part_sq4(flip_sign(b0, sgn)) + pos_part_sq4(flip_sign(f0, sgn));
        s4 = (s4 + neg_part_sq4(flip_sign(b1, sgn))) + pos_part_sq4(flip_sign(f1, sgn));
        s4 = (s4 + neg_part_sq4(flip_sign(b2, sgn))) + pos_part_sq4(flip_sign(f2, sgn));
        T r = T(0.5) * sqrt_nonneg(s4);
        if (!scaled_sum_ok(s4)) r = gmag_os_exact<T>(b0, f0, b1, f1, b2, f2, neg);
        res[v] = r;
    }
}

// centered differences of the N voxels (masked_gradient.c:24-71): 0.5 (plus - minus), one-sided
// at the borders, / dx, magnitude, optional normalisation
template <typename T, bool EXACT, bool UNIT, bool BORDER>
__device__ __forceinline__ void tma_centered(const TmaStencil<T>& q, const StencilParams& p, bool ilo,
                                             bool ihi, bool jlo, bool jhi, bool klo, bool khi,
                                             int normalize, T (&d0)[VecOf<T>::N], T (&d1)[VecOf<T>::N],
                                             T (&d2)[VecOf<T>::N], T (&g)[VecOf<T>::N]) {
    constexpr int N = VecOf<T>::N;
#pragma unroll
    for (int v = 0; v < N; ++v) {
        const T cc = q.cur[v];
        const T rp = v + 1 < N ? q.cur[v + 1 < N ? v + 1 : v] : q.right;
        const T rm = v > 0 ? q.cur[v > 0 ? v - 1 : 0] : q.left;
        T a0 = T(0.5) * (q.nxt[v] - q.prv[v]), a1 = T(0.5) * (q.jp[v] - q.jm[v]), a2 = T(0.5) * (rp - rm);
        if (BORDER) {
            if (ilo) a0 = q.nxt[v] - cc; else if (ihi) a0 = cc - q.prv[v];
            if (jlo) a1 = q.jp[v] - cc; else if (jhi) a1 = cc - q.jm[v];
        }
        if (v == 0 && klo) a2 = rp - cc;
        if (v == N - 1 && khi) a2 = cc - rm;
        if (!UNIT) {
            a0 = scale_dx<T, EXACT>(a0, (T)p.dx0, (T)p.r0);
            a1 = scale_dx<T, EXACT>(a1, (T)p.dx1, (T)p.r1);
            a2 = scale_dx<T, EXACT>(a2, (T)p.dx2, (T)p.r2);
        }
        const T gg = dev_sqrt((a0 * a0 + a1 * a1) + a2 * a2);
        g[v] = gg;
        if (normalize == 1 && gg > T(0)) { a0 /= gg; a1 /= gg; a2 /= gg; }
        d0[v] = a0; d1[v] = a1; d2[v] = a2;
    }
}

template <typename T>
struct TmaOut {
    T* out;              // OP 0: |Du|; OP 1: gmag
    T* grads;            // OP 1: three arrays `plane` elements apart
    i64 plane;
    int normalize;
};

// OP 0: Osher-Sethian upwind |Du| (reads nu); OP 1: centered gradient + magnitude (no nu)
template <typename T, int OP, bool EXACT, bool UNIT>
__global__ void __launch_bounds__(TM_THREADS)
gmag_os_tma_kernel(const __grid_constant__ TmaMaps maps, int has_mask, TmaOut<T> o,
                   StencilParams p, int chunk /* planes per CTA */) {
    typedef typename VecOf<T>::type V;
    constexpr int N = VecOf<T>::N;                       // voxels per thread
    constexpr int TK = TM_ROW_BYTES / (int)sizeof(T);    // voxels per tile row
    constexpr int M_BYTES = TM_TJ * TK;                  // mask tile
    T* __restrict__ out = o.out;
    extern __shared__ __align__(128) unsigned char tm_smem[];
    // TMA destinations must be 128-byte aligned
    unsigned char* sA = tm_smem + ((128u - (smem_u32(tm_smem) & 127u)) & 127u);   // A_RING A tiles
    unsigned char* sNu = sA + TM_A_RING * TM_A_SLOT;                     // STAGES nu tiles
    unsigned char* sM = sNu + TM_STAGES * TM_NU_BYTES;                   // STAGES mask tiles
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(sM + TM_STAGES * TM_M_MAX);
    // bars[0 .. STAGES): full, bars[STAGES .. 2 STAGES): empty, bars[2 STAGES]: prologue
    const unsigned bar0 = smem_u32(bars);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // launch order: j tiles fastest, then k tiles, then (item, chunk): CTAs that are resident
    // together are neighbours in j and k, so their halo rows / columns hit L2
    const unsigned chunks = (unsigned)((p.n0 + chunk - 1) / chunk);
    const int item = (int)(blockIdx.z / chunks), ci = (int)(blockIdx.z % chunks);
    const int i0 = ci * chunk, i1 = min(p.n0, i0 + chunk), L = i1 - i0;
    const int j0 = blockIdx.x * TM_TJ, k0 = blockIdx.y * TK;
    if (threadIdx.x == 0) {
        for (int s = 0; s < TM_STAGES; ++s) {
            mbar_init(bar0 + 8 * s, 1);                          // full: the producer's expect_tx
            mbar_init(bar0 + 8 * (TM_STAGES + s), TM_TJ);        // empty: one arrive per consumer warp
        }
        mbar_init(bar0 + 8 * 2 * TM_STAGES, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned stage_bytes = TM_A_BYTES + (OP == 0 ? TM_NU_BYTES : 0) + (has_mask ? M_BYTES : 0);
    if (warp == TM_TJ) {
        // ---------------------------------------------------------------- producer warp
        if (lane == 0) {
            // prologue: A planes i0-1 and i0 (ring slots 0 and 1); the A box starts N voxels (16
            // bytes) left of the tile and one row above it
            const unsigned pro = bar0 + 8 * 2 * TM_STAGES;
            mbar_expect_tx(pro, 2 * TM_A_BYTES);
            tma_load_4d(smem_u32(sA), &maps.a, pro, k0 - N, j0 - 1, i0 - 1, item);
            tma_load_4d(smem_u32(sA + TM_A_SLOT), &maps.a, pro, k0 - N, j0 - 1, i0, item);
            for (int u = 0; u < L; ++u) {
                const int s = u % TM_STAGES;
                if (u >= TM_STAGES) mbar_wait(bar0 + 8 * (TM_STAGES + s), ((u / TM_STAGES) - 1) & 1);
                const unsigned full = bar0 + 8 * s;
                mbar_expect_tx(full, stage_bytes);
                // A plane i0+u+1 -> ring slot (u+2) % A_RING (the slot of plane i0+u-STAGES-1, last
                // read in iteration u-STAGES-1); nu / mask plane i0+u -> slot s
                tma_load_4d(smem_u32(sA + ((u + 2) % TM_A_RING) * TM_A_SLOT), &maps.a, full,
                            k0 - N, j0 - 1, i0 + u + 1, item);
                if (OP == 0)
                    tma_load_4d(smem_u32(sNu + s * TM_NU_BYTES), &maps.nu, full, k0, j0, i0 + u, item);
                if (has_mask)
                    tma_load_4d(smem_u32(sM + s * TM_M_MAX), &maps.m, full, k0, j0, i0 + u, item);
            }
        }
        return;
    }
    // -------------------------------------------------------------------- consumer warps
    const int ty = warp, tx = lane;
    const int j = j0 + ty, k = k0 + N * tx;
    const bool inside = j < p.n1 && k < p.n2;                 // (n2 % N == 0: all N voxels or none)
    const bool jlo = j == 0, jhi = j == p.n1 - 1, klo = k == 0, khi = k + N == p.n2;
    // byte offset of this thread's centre vector inside an A tile / a nu tile / a mask tile
    const unsigned oA = (ty + 1) * TM_AROW_BYTES + 16 * tx + 16;
    const unsigned oNu = ty * TM_ROW_BYTES + 16 * tx;
    const unsigned oM = ty * TK + N * tx;
    T* po = out + (((i64)item * p.n0 + i0) * p.n1 + j) * (i64)p.n2 + k;
    mbar_wait(bar0 + 8 * 2 * TM_STAGES, 0);
    TmaStencil<T> q;
    unpack(*reinterpret_cast<const V*>(sA + oA), q.prv);
    unpack(*reinterpret_cast<const V*>(sA + TM_A_SLOT + oA), q.cur);
    for (int t = 0; t < L; ++t) {
        const int s = t % TM_STAGES;
        mbar_wait(bar0 + 8 * s, (t / TM_STAGES) & 1);
        const unsigned char* tc = sA + ((t + 1) % TM_A_RING) * TM_A_SLOT;      // plane i
        const unsigned char* tn = sA + ((t + 2) % TM_A_RING) * TM_A_SLOT;      // plane i+1
        unpack(*reinterpret_cast<const V*>(tn + oA), q.nxt);
        unsigned m = (1u << N) - 1u;
        if (has_mask) {
            unsigned raw;
            if (N == 2) raw = *reinterpret_cast<const unsigned short*>(sM + s * TM_M_MAX + oM);
            else raw = *reinterpret_cast<const unsigned*>(sM + s * TM_M_MAX + oM);
            m = 0u;
#pragma unroll
            for (int v = 0; v < N; ++v) m |= ((raw >> (8 * v)) & 0xffu) ? (1u << v) : 0u;
        }
        if (m && inside) {
            unpack(*reinterpret_cast<const V*>(tc + oA + TM_AROW_BYTES), q.jp);
            unpack(*reinterpret_cast<const V*>(tc + oA - TM_AROW_BYTES), q.jm);
            q.left = *reinterpret_cast<const T*>(tc + oA - sizeof(T));
            q.right = *reinterpret_cast<const T*>(tc + oA + 16);
            const int i = i0 + t;
            const bool ilo = i == 0, ihi = i == p.n0 - 1;
            // Border rows (j = 0 / n1-1: one warp of a border CTA) and border planes (i = 0 /
            // n0-1: CTA-uniform) take the general path; everything else only keeps the two
            // k-border selects (first / last thread of the first / last k tile).  The split keeps
            // ~30 select instructions per iteration out of the common path.
            if (OP == 0) {
                T nuv[N], res[N];
                unpack(*reinterpret_cast<const V*>(sNu + s * TM_NU_BYTES + oNu), nuv);
                if (ilo | ihi | jlo | jhi)
                    tma_upwind<T, EXACT, UNIT, true>(q, nuv, p, ilo, ihi, jlo, jhi, klo, khi, res);
                else
                    tma_upwind<T, EXACT, UNIT, false>(q, nuv, p, false, false, false, false, klo, khi, res);
                if (m == (1u << N) - 1u) {
                    *reinterpret_cast<V*>(po) = pack(res);
                } else {
#pragma unroll
                    for (int v = 0; v < N; ++v)
                        if ((m >> v) & 1u) po[v] = res[v];
                }
            } else {
                T d0[N], d1[N], d2[N], g[N];
                if (ilo | ihi | jlo | jhi)
                    tma_centered<T, EXACT, UNIT, true>(q, p, ilo, ihi, jlo, jhi, klo, khi, o.normalize,
                                                       d0, d1, d2, g);
                else
                    tma_centered<T, EXACT, UNIT, false>(q, p, false, false, false, false, klo, khi,
                                                        o.normalize, d0, d1, d2, g);
                T* g0 = o.grads + (po - out);
                if (m == (1u << N) - 1u) {
                    *reinterpret_cast<V*>(po) = pack(g);
                    *reinterpret_cast<V*>(g0) = pack(d0);
                    *reinterpret_cast<V*>(g0 + o.plane) = pack(d1);
                    *reinterpret_cast<V*>(g0 + 2 * o.plane) = pack(d2);
                } else {
#pragma unroll
                    for (int v = 0; v < N; ++v)
                        if ((m >> v) & 1u) {
                            po[v] = g[v]; g0[v] = d0[v]; g0[o.plane + v] = d1[v]; g0[2 * o.plane + v] = d2[v];
                        }
                }
            }
        }
        // this warp is done with stage s (its nu / mask tile, and the A plane i-1 behind it)
        __syncwarp();
        if (lane == 0) mbar_arrive(bar0 + 8 * (TM_STAGES + s));
#pragma unroll
        for (int v = 0; v < N; ++v) { q.prv[v] = q.cur[v]; q.cur[v] = q.nxt[v]; }
        po += p.st0;
    }
}

static bool marchv_eligible(const Geom& g, const void* a, const void* b, const void* c,
                            const void* m, int vec, size_t elem) {
    auto al = [&](const void* q, size_t n) { return q == nullptr || ((size_t)q % n) == 0; };
    return g.ndim == 3 && g.n[2] % vec == 0 && g.n[2] >= 32 * vec && g.st[0] < ((i64)1 << 30) &&
           al(a, vec * elem) && al(b, vec * elem) && al(c, vec * elem) && al(m, vec);
}

static void marchv_launch_shape(const Geom& g, int vec, dim3& grid, dim3& block) {
    const i64 chunks = (g.n[0] + MV_CHUNK - 1) / MV_CHUNK;
    block = dim3(MV_BX, MV_BY, 1);
    grid = dim3((unsigned)(g.batch * chunks), (unsigned)((g.n[1] + MV_BY - 1) / MV_BY),
                (unsigned)((g.n[2] + MV_BX * vec - 1) / (MV_BX * vec)));
}

static bool march_eligible(const Geom& g) {
    // 32-bit in-chunk offsets: (chunk + 1) planes must stay below 2^31 elements
    return g.ndim == 3 && (i64)(MARCH_CHUNK + 2) * g.st[0] < ((i64)1 << 31) && g.n[2] >= 32;
}

static void march_launch_shape(const Geom& g, dim3& grid, dim3& block) {
    const i64 chunks = (g.n[0] + MARCH_CHUNK - 1) / MARCH_CHUNK;
    block = dim3(MARCH_BX, MARCH_BY, 1);
    grid = dim3((unsigned)(g.batch * chunks), (unsigned)((g.n[1] + MARCH_BY - 1) / MARCH_BY),
                (unsigned)((g.n[2] + MARCH_BX - 1) / MARCH_BX));
}

static void launch_shape(const Geom& g, dim3& grid, dim3& block, StencilParams& p) {
    int bx = 32;
    while (bx < 256 && bx < g.n[2]) bx <<= 1;
    int by = 256 / bx;
    p.n0 = g.n[0]; p.n1 = g.n[1]; p.n2 = g.n[2];
    p.st0 = g.st[0]; p.st1 = g.st[1];
    p.rows = g.batch * g.n[0] * g.n[1];
    p.dx0 = g.dx[0]; p.dx1 = g.dx[1]; p.dx2 = g.dx[2];
    p.r0 = g.rdx[0]; p.r1 = g.rdx[1]; p.r2 = g.rdx[2];
    if (by > g.n[1]) { by = 1; while (by < g.n[1]) by <<= 1; }
    block = dim3(bx, by, 1);
    grid = dim3((unsigned)(g.batch * g.n[0]), (unsigned)((g.n[1] + by - 1) / by),
                (unsigned)((g.n[2] + bx - 1) / bx));
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda at link time)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            (void)cudaGetLastError();
    }
    return fn;
}

// tensor map over (k, j, i, item) of a dense (batch, n0, n1, n2) array with `elem`-byte elements
static bool make_map(void* out, CUtensorMapDataType type, size_t elem, const void* base,
                     const Geom& g, unsigned box_k, unsigned box_j) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (fn == nullptr) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)g.n[2], (cuuint64_t)g.n[1], (cuuint64_t)g.n[0],
                                (cuuint64_t)g.batch};
    const cuuint64_t strides[3] = {(cuuint64_t)g.n[2] * elem, (cuuint64_t)g.st[0] * elem,
                                   (cuuint64_t)g.voxels * elem};
    const cuuint32_t box[4] = {box_k, box_j, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return fn(reinterpret_cast<CUtensorMap*>(out), type, 4, const_cast<void*>(base), dims, strides,
              box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// A/B switch (probes and tests cover both generations of the kernel): lsml_stencil_use_tma()
static bool g_stencil_tma = getenv("LSML_STENCIL_NO_TMA") == nullptr;
int stencil_use_tma(int on) {
    const int prev = g_stencil_tma ? 1 : 0;
    if (on >= 0) g_stencil_tma = on != 0;
    return prev;
}

// 3-D grids whose rows are 16-byte multiples: the TMA plane pipeline.  *done = false when the
// grid is not eligible (the caller falls back to the register-marching kernels).
// nu == nullptr selects the centered gradient (grads / plane / normalize are then used).
template <typename T>
static int stencil_tma(const Geom& g, const T* A, const u8* mask, const T* nu, T* out, T* grads,
                       i64 plane, int normalize, const StencilParams& p, bool exact,
                       cudaStream_t s, bool* done) {
    constexpr int N = VecOf<T>::N;
    constexpr int TK = TM_ROW_BYTES / (int)sizeof(T);
    const CUtensorMapDataType dtype = sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64
                                                     : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    *done = false;
    const bool off = !g_stencil_tma;
    auto al16 = [](const void* q) { return ((size_t)q & 15) == 0; };
    const bool centered = nu == nullptr;
    if (off || g.ndim != 3 || g.n[2] % N != 0 || g.n[2] < TK || g.n[1] < 8 || g.n[0] < 2 ||
        !al16(A) || !al16(nu) || !al16(out) || (mask && !al16(mask)) || (mask && g.n[2] % 16 != 0) ||
        (centered && (!al16(grads) || plane % N != 0)))
        return 0;
    TmaMaps maps;
    if (!make_map(maps.a, dtype, sizeof(T), A, g, TK + 2 * N, TM_TJ + 2) ||
        (!centered && !make_map(maps.nu, dtype, sizeof(T), nu, g, TK, TM_TJ)) ||
        (mask && !make_map(maps.m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, mask, g, TK, TM_TJ)))
        return 0;
    const bool unit = g.dx[0] == 1.0 && g.dx[1] == 1.0 && g.dx[2] == 1.0;
    // Planes per CTA: every CTA pays two extra A planes (its prologue), and the grid should fill
    // whole waves of the 3 CTAs/SM that fit (72 registers x 288 threads): pick the balanced chunk
    // length with the smallest  waves x (planes + 2).
    const i64 tiles = (i64)((g.n[1] + TM_TJ - 1) / TM_TJ) * ((g.n[2] + TK - 1) / TK) * g.batch;
    const i64 slots = (i64)LSML_SMS * TM_MIN_CTAS;
    int chunk = g.n[0];
    i64 best = -1;
    for (int c = 1; c <= g.n[0]; ++c) {
        const int len = (g.n[0] + c - 1) / c;
        if (len < 8 && c > 1) break;
        const i64 ctas = tiles * ((g.n[0] + len - 1) / len);
        const i64 cost = ((ctas + slots - 1) / slots) * (len + 2);
        if (best < 0 || cost < best) { best = cost; chunk = len; }
    }
    if (const char* e = getenv("LSML_TMA_CHUNK")) { const int v = atoi(e); if (v >= 1) chunk = v < g.n[0] ? v : g.n[0]; }
    const i64 chunks = (g.n[0] + chunk - 1) / chunk;
    if (g.batch * chunks > 65535) return 0;
    const dim3 grid((unsigned)((g.n[1] + TM_TJ - 1) / TM_TJ), (unsigned)((g.n[2] + TK - 1) / TK),
                    (unsigned)(g.batch * chunks));
    static bool configured = false;      // (one flag per element type: function-local static of a template)
    if (!configured) {
#define LSML_TMA_ATTR(K) LSML_CUDA(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TM_SMEM))
        LSML_TMA_ATTR((gmag_os_tma_kernel<T, 0, true, true>)); LSML_TMA_ATTR((gmag_os_tma_kernel<T, 0, true, false>));
        LSML_TMA_ATTR((gmag_os_tma_kernel<T, 0, false, false>)); LSML_TMA_ATTR((gmag_os_tma_kernel<T, 1, true, true>));
        LSML_TMA_ATTR((gmag_os_tma_kernel<T, 1, true, false>)); LSML_TMA_ATTR((gmag_os_tma_kernel<T, 1, false, false>));
#undef LSML_TMA_ATTR
        configured = true;
    }
    const int has_mask = mask != nullptr;
    const TmaOut<T> o{out, grads, plane, normalize};
#define LSML_TMA_GO(OP)                                                                                         \
    do {                                                                                                         \
        if (unit) gmag_os_tma_kernel<T, OP, true, true><<<grid, TM_THREADS, TM_SMEM, s>>>(maps, has_mask, o, p, chunk);        \
        else if (exact) gmag_os_tma_kernel<T, OP, true, false><<<grid, TM_THREADS, TM_SMEM, s>>>(maps, has_mask, o, p, chunk); \
        else gmag_os_tma_kernel<T, OP, false, false><<<grid, TM_THREADS, TM_SMEM, s>>>(maps, has_mask, o, p, chunk);           \
    } while (0)
    if (centered) LSML_TMA_GO(1); else LSML_TMA_GO(0);
#undef LSML_TMA_GO
    LSML_LAUNCH_CHECK("gmag_os_tma_kernel");
    *done = true;
    return 0;
}
template <typename T>
static int gmag_os_tma(const Geom& g, const T* A, const u8* mask, const T* nu, T* out,
                       const StencilParams& p, bool exact, cudaStream_t s, bool* done) {
    return stencil_tma<T>(g, A, mask, nu, out, nullptr, 0, 0, p, exact, s, done);
}
template <typename T>
static int centered_tma(const Geom& g, const T* A, const u8* mask, T* grads, T* gmag, i64 plane,
                        int normalize, const StencilParams& p, bool exact, cudaStream_t s,
                        bool* done) {
    return stencil_tma<T>(g, A, mask, nullptr, gmag, grads, plane, normalize, p, exact, s, done);
}

template <typename T>
int gmag_os_dense(const Geom& g, const T* A, const u8* mask, const T* nu, T* out,
                  cudaStream_t s) {
    if (g.voxels * g.batch == 0) return 0;
    dim3 grid, block; StencilParams p;
    launch_shape(g, grid, block, p);
    const bool exact = g.rdx[0] != 0.0 && g.rdx[1] != 0.0 && g.rdx[2] != 0.0;
    if (g.st[0] >= (i64)1 << 31) { set_error("grid plane too large for 32-bit strides"); return 1; }
    {
        bool done = false;
        const int st = gmag_os_tma(g, A, mask, nu, out, p, exact, s, &done);
        if (st || done) return st;
    }
    if (marchv_eligible(g, A, nu, out, mask, VecOf<T>::N, sizeof(T))) {
        const bool unit = g.dx[0] == 1.0 && g.dx[1] == 1.0 && g.dx[2] == 1.0;
        dim3 mg, mb;
        marchv_launch_shape(g, VecOf<T>::N, mg, mb);
        if (unit) gmag_os_marchv_kernel<T, true, true><<<mg, mb, 0, s>>>(A, mask, nu, out, p);
        else if (exact) gmag_os_marchv_kernel<T, true, false><<<mg, mb, 0, s>>>(A, mask, nu, out, p);
        else gmag_os_marchv_kernel<T, false, false><<<mg, mb, 0, s>>>(A, mask, nu, out, p);
        LSML_LAUNCH_CHECK("gmag_os_marchv_kernel");
        return 0;
    }
    if (march_eligible(g)) {
        const bool unit = g.dx[0] == 1.0 && g.dx[1] == 1.0 && g.dx[2] == 1.0;
        dim3 mg, mb;
        march_launch_shape(g, mg, mb);
        if (unit) gmag_os_march3d_kernel<T, true, true><<<mg, mb, 0, s>>>(A, mask, nu, out, p);
        else if (exact) gmag_os_march3d_kernel<T, true, false><<<mg, mb, 0, s>>>(A, mask, nu, out, p);
        else gmag_os_march3d_kernel<T, false, false><<<mg, mb, 0, s>>>(A, mask, nu, out, p);
        LSML_LAUNCH_CHECK("gmag_os_march3d_kernel");
        return 0;
    }
#define LSML_GO(ND)                                                                        \
    do {                                                                                   \
        if (exact) gmag_os_kernel<T, ND, true><<<grid, block, 0, s>>>(A, mask, nu, out, p); \
        else gmag_os_kernel<T, ND, false><<<grid, block, 0, s>>>(A, mask, nu, out, p);      \
    } while (0)
    if (g.ndim == 3) LSML_GO(3); else if (g.ndim == 2) LSML_GO(2); else LSML_GO(1);
#undef LSML_GO
    LSML_LAUNCH_CHECK("gmag_os_kernel");
    return 0;
}

template <typename T>
int gradient_centered_dense(const Geom& g, const T* A, const u8* mask, T* grads, T* gmag,
                            int normalize, cudaStream_t s) {
    if (g.voxels * g.batch == 0) return 0;
    dim3 grid, block; StencilParams p;
    launch_shape(g, grid, block, p);
    const i64 plane = g.voxels * g.batch;
    const bool exact = g.rdx[0] != 0.0 && g.rdx[1] != 0.0 && g.rdx[2] != 0.0;
    if (g.st[0] >= (i64)1 << 31) { set_error("grid plane too large for 32-bit strides"); return 1; }
    {
        bool done = false;
        const int st = centered_tma(g, A, mask, grads, gmag, plane, normalize, p, exact, s, &done);
        if (st || done) return st;
    }
    if (march_eligible(g)) {
        const bool unit = g.dx[0] == 1.0 && g.dx[1] == 1.0 && g.dx[2] == 1.0;
        dim3 mg, mb;
        march_launch_shape(g, mg, mb);
        if (unit) gradient_centered_march3d_kernel<T, true, true><<<mg, mb, 0, s>>>(A, mask, grads, gmag, p, plane, normalize);
        else if (exact) gradient_centered_march3d_kernel<T, true, false><<<mg, mb, 0, s>>>(A, mask, grads, gmag, p, plane, normalize);
        else gradient_centered_march3d_kernel<T, false, false><<<mg, mb, 0, s>>>(A, mask, grads, gmag, p, plane, normalize);
        LSML_LAUNCH_CHECK("gradient_centered_march3d_kernel");
        return 0;
    }
#define LSML_GC(ND)                                                                               \
    do {                                                                                          \
        if (exact) gradient_centered_kernel<T, ND, true><<<grid, block, 0, s>>>(A, mask, grads, gmag, p, plane, normalize); \
        else gradient_centered_kernel<T, ND, false><<<grid, block, 0, s>>>(A, mask, grads, gmag, p, plane, normalize);      \
    } while (0)
    if (g.ndim == 3) LSML_GC(3); else if (g.ndim == 2) LSML_GC(2); else LSML_GC(1);
#undef LSML_GC
    LSML_LAUNCH_CHECK("gradient_centered_kernel");
    return 0;
}

template int gmag_os_dense<double>(const Geom&, const double*, const u8*, const double*, double*, cudaStream_t);
template int gmag_os_dense<float>(const Geom&, const float*, const u8*, const float*, float*, cudaStream_t);
template int gradient_centered_dense<double>(const Geom&, const double*, const u8*, double*, double*, int, cudaStream_t);
template int gradient_centered_dense<float>(const Geom&, const float*, const u8*, float*, float*, int, cudaStream_t);

}  // namespace lsml
