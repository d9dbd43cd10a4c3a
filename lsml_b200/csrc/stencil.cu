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

// one-sided / two-sided differences along one axis (masked_gradient.c:165-176)
template <typename T>
__device__ __forceinline__ void fwd_bwd(const T* __restrict__ A, i64 l, i64 st, int c, int n,
                                        T center, T& f, T& b) {
    if (n == 1) { f = T(0); b = T(0); return; }   // the reference reads out of bounds here
    if (c == 0) { f = __ldg(A + l + st) - center; b = f; }
    else if (c == n - 1) { b = center - __ldg(A + l - st); f = b; }
    else { f = __ldg(A + l + st) - center; b = center - __ldg(A + l - st); }
}

template <typename T>
__device__ __forceinline__ T os_terms(T b, T f, bool neg, T acc, bool first) {
    T tb, tf;
    if (neg) { tb = ref_max(b, T(0)); tf = ref_min(f, T(0)); }
    else     { tb = ref_min(b, T(0)); tf = ref_max(f, T(0)); }
    tb = tb * tb; tf = tf * tf;
    return first ? (tb + tf) : ((acc + tb) + tf);
}

template <typename T, int NDIM>
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
    const T c = __ldg(A + l);
    const bool neg = __ldg(nu + l) < T(0);
    T acc = T(0), f, b;
    bool first = true;
    if (NDIM >= 3) {
        fwd_bwd(A, l, p.st0, i, p.n0, c, f, b);
        f = div_dx(f, (T)p.dx0, (T)p.r0); b = div_dx(b, (T)p.dx0, (T)p.r0);
        acc = os_terms(b, f, neg, acc, first); first = false;
    }
    if (NDIM >= 2) {
        fwd_bwd(A, l, p.st1, j, p.n1, c, f, b);
        f = div_dx(f, (T)p.dx1, (T)p.r1); b = div_dx(b, (T)p.dx1, (T)p.r1);
        acc = os_terms(b, f, neg, acc, first); first = false;
    }
    fwd_bwd(A, l, (i64)1, k, p.n2, c, f, b);
    f = div_dx(f, (T)p.dx2, (T)p.r2); b = div_dx(b, (T)p.dx2, (T)p.r2);
    acc = os_terms(b, f, neg, acc, first);
    out[l] = dev_sqrt(acc);
}

// centered difference along one axis (masked_gradient.c:24-60)
template <typename T>
__device__ __forceinline__ T centered(const T* __restrict__ A, i64 l, i64 st, int c, int n,
                                      T center) {
    if (n == 1) return T(0);
    if (c == 0) return __ldg(A + l + st) - center;
    if (c == n - 1) return center - __ldg(A + l - st);
    return T(0.5) * (__ldg(A + l + st) - __ldg(A + l - st));
}

template <typename T, int NDIM>
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
    T d0 = T(0), d1 = T(0), d2;
    if (NDIM >= 3) {
        d0 = div_dx(centered(A, l, p.st0, i, p.n0, c), (T)p.dx0, (T)p.r0);
    }
    if (NDIM >= 2) {
        d1 = div_dx(centered(A, l, p.st1, j, p.n1, c), (T)p.dx1, (T)p.r1);
    }
    d2 = div_dx(centered(A, l, (i64)1, k, p.n2, c), (T)p.dx2, (T)p.r2);
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
    if (g.ndim == 3) gmag_os_kernel<T, 3><<<grid, block, 0, s>>>(A, mask, nu, out, p);
    else if (g.ndim == 2) gmag_os_kernel<T, 2><<<grid, block, 0, s>>>(A, mask, nu, out, p);
    else gmag_os_kernel<T, 1><<<grid, block, 0, s>>>(A, mask, nu, out, p);
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
    if (g.ndim == 3)
        gradient_centered_kernel<T, 3><<<grid, block, 0, s>>>(A, mask, grads, gmag, p, plane, normalize);
    else if (g.ndim == 2)
        gradient_centered_kernel<T, 2><<<grid, block, 0, s>>>(A, mask, grads, gmag, p, plane, normalize);
    else
        gradient_centered_kernel<T, 1><<<grid, block, 0, s>>>(A, mask, grads, gmag, p, plane, normalize);
    LSML_LAUNCH_CHECK("gradient_centered_kernel");
    return 0;
}

template int gmag_os_dense<double>(const Geom&, const double*, const u8*, const double*, double*, cudaStream_t);
template int gmag_os_dense<float>(const Geom&, const float*, const u8*, const float*, float*, cudaStream_t);
template int gradient_centered_dense<double>(const Geom&, const double*, const u8*, double*, double*, int, cudaStream_t);
template int gradient_centered_dense<float>(const Geom&, const float*, const u8*, float*, float*, int, cudaStream_t);

}  // namespace lsml
