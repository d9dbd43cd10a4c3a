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

// sqrt(0) == 0 without entering the library's slow path (local extrema of u give exactly 0)
template <typename T>
__device__ __forceinline__ T sqrt_nonneg(T x) { return x > T(0) ? dev_sqrt(x) : x; }

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

template <typename T>
int gmag_os_dense(const Geom& g, const T* A, const u8* mask, const T* nu, T* out,
                  cudaStream_t s) {
    if (g.voxels * g.batch == 0) return 0;
    dim3 grid, block; StencilParams p;
    launch_shape(g, grid, block, p);
    const bool exact = g.rdx[0] != 0.0 && g.rdx[1] != 0.0 && g.rdx[2] != 0.0;
    if (g.st[0] >= (i64)1 << 31) { set_error("grid plane too large for 32-bit strides"); return 1; }
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
