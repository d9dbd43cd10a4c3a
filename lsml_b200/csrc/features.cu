// features.cu -- per-voxel feature extraction on the narrow band.
//
// One-time per image (u-independent, lsml/feature/provided/image.py):
//   gauss1d_kernel        scipy.ndimage.gaussian_filter1d (order 0/1, mode='reflect'), evaluated
//                         in scipy's NI_Correlate1D order so the planes are bit-identical:
//                         tmp = x[0]*w[0]; for jj=-r..-1: tmp += (x[jj] +- x[-jj]) * w[jj]
//   np_gradient_mag_kernel numpy.gradient(img, *dx) magnitude (ImageEdgeSample sigma=0)
//   normalize_kernel      (img - mean) / std per item (lsml/core/model.py:328-329)
// Every iteration:
//   global_stats_kernel   {u>0}: count, sum idx, sum idx^2 per axis as EXACT int64 (Size,
//                         Moments, centre of mass; shape.py:28, 214-230)
//   band_sum_kernel       per-item band sums of the image planes (InteriorImageAverage /
//                         Variation are band statistics, image.py:83-119) -- two-stage,
//                         fixed-shape, deterministic
//   contour_length_kernel marching squares total length (BoundarySize 2-D, shape.py:70-85)
//   item_features_kernel  turns the raw sums into the per-item ("global") feature values
//   assemble_kernel       writes the (B, F) feature matrix in band order (feature_map.py:53-80)
#include <cstdlib>
#include "lsml_internal.h"
#include "band_features.cuh"

namespace lsml {

// ---------------------------------------------------------------------------------------------
// separable Gaussian (order 0: symmetric, order 1: antisymmetric), scipy 'reflect' boundary
// ---------------------------------------------------------------------------------------------
enum GaussMode { GM_STORE = 0, GM_SQ_FIRST = 1, GM_SQ_ADD = 2, GM_SQ_ADD_SQRT = 3, GM_SQ_SQRT = 4 };

__device__ __forceinline__ int reflect_index(int i, int n) {
    // (d c b a | a b c d | d c b a): period 2n
    const int period = 2 * n;
    int m = i % period;
    if (m < 0) m += period;
    return m < n ? m : period - 1 - m;
}

__global__ void __launch_bounds__(256)
gauss1d_kernel(const double* __restrict__ in, double* __restrict__ out,
               const double* __restrict__ w /* centre at w[r] */, int r, int antisym,
               int n0, int n1, int n2, int axis, i64 rows, int mode) {
    // blockIdx.x = item*n0 + i, blockIdx.y = j tile, blockIdx.z = k tile (no per-thread division)
    const int k = blockIdx.z * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (k >= n2 || j >= n1) return;
    const i64 l = ((i64)blockIdx.x * n1 + j) * n2 + k;
    int c, n; i64 st;
    if (axis == 2) { c = k; n = n2; st = 1; }
    else if (axis == 1) { c = j; n = n1; st = n2; }
    else { c = (int)(blockIdx.x % (unsigned)n0); n = n0; st = (i64)n1 * n2; }
    const double* base = in + (l - (i64)c * st);
    double tmp = __ldg(base + (i64)c * st) * __ldg(w + r);
    const bool interior = (c - r >= 0) && (c + r < n);
    for (int jj = -r; jj < 0; ++jj) {
        const int ia = interior ? c + jj : reflect_index(c + jj, n);
        const int ib = interior ? c - jj : reflect_index(c - jj, n);
        const double a = __ldg(base + (i64)ia * st), b = __ldg(base + (i64)ib * st);
        const double pair = antisym ? (a - b) : (a + b);
        tmp = tmp + pair * __ldg(w + r + jj);
    }
    if (mode == GM_STORE) out[l] = tmp;
    else if (mode == GM_SQ_FIRST) out[l] = tmp * tmp;
    else if (mode == GM_SQ_ADD) out[l] = out[l] + tmp * tmp;
    else if (mode == GM_SQ_ADD_SQRT) out[l] = sqrt(out[l] + tmp * tmp);
    else out[l] = sqrt(tmp * tmp);
}

// The same filter along a STRIDED axis (0 or 1) as a register march: a thread owns one column,
// walks it from end to end and keeps the 2 R + 1 taps of the current voxel (plus GM_AHEAD loads in
// flight) in a register ring, so every input element is read ONCE, by neighbouring lanes in
// neighbouring addresses.  gauss1d_kernel re-reads each element 2 R + 1 times through L1 with
// little reuse across a CTA's few rows: 250 us per pass at 256^3 (16 % of the HBM peak).  Same
// expression, same order (scipy's NI_Correlate1D: centre first, then the pairs outside in), hence
// the same bits.  The ring is indexed with compile-time slots: the march is unrolled over one
// ring length.
constexpr int GM_AHEAD = 8;
constexpr int GM_THREADS = 128;

template <int R>
__global__ void __launch_bounds__(GM_THREADS)
gauss1d_march_kernel(const double* __restrict__ in, double* __restrict__ out,
                     const double* __restrict__ w /* centre at w[R] */, int antisym, int n, i64 st,
                     i64 inner_count, i64 outer_stride, int mode) {
    constexpr int W = 2 * R + 1 + GM_AHEAD;
    const i64 q = (i64)blockIdx.y * GM_THREADS + threadIdx.x;
    if (q >= inner_count) return;
    const i64 base = (i64)blockIdx.x * outer_stride + q;
    const double* col = in + base;
    double wt[R + 1];
#pragma unroll
    for (int k = 0; k <= R; ++k) wt[k] = __ldg(w + k);
    auto tap = [&](int p) -> double {              // element p of the column, 'reflect' boundary
        const int pi = (unsigned)p < (unsigned)n ? p : reflect_index(p, n);
        return __ldg(col + (i64)pi * st);
    };
    double win[W];                                 // slot of element p: (p + R) mod W
#pragma unroll
    for (int p = -R; p < R + GM_AHEAD; ++p) win[(p + R) % W] = tap(p);
    for (int c0 = 0; c0 < n; c0 += W) {
#pragma unroll
        for (int s = 0; s < W; ++s) {
            const int c = c0 + s;
            if (c >= n) break;
            win[(s + 2 * R + GM_AHEAD) % W] = tap(c + R + GM_AHEAD);
            double tmp = win[(s + R) % W] * wt[R];
#pragma unroll
            for (int jj = -R; jj < 0; ++jj) {
                const double a = win[(s + jj + R) % W], b = win[(s - jj + R) % W];
                const double pair = antisym ? (a - b) : (a + b);
                tmp = tmp + pair * wt[R + jj];
            }
            const i64 l = base + (i64)c * st;
            if (mode == GM_STORE) out[l] = tmp;
            else if (mode == GM_SQ_FIRST) out[l] = tmp * tmp;
            else if (mode == GM_SQ_ADD) out[l] = out[l] + tmp * tmp;
            else if (mode == GM_SQ_ADD_SQRT) out[l] = sqrt(out[l] + tmp * tmp);
            else out[l] = sqrt(tmp * tmp);
        }
    }
}

// Both passes of scipy.ndimage.gaussian_filter on a batch of SMALL 2-D items in one kernel: one
// CTA per item, the item lives in shared memory (the axis-0 pass row by row), so a plane
// costs one global read and one global write per pixel instead of two of each, and every tap is
// a conflict-free shared-memory load.  Same expressions in the same order as gauss1d_kernel
// (axis 0 first, then axis 1; symmetric pairing), hence bit-identical.  Used for the level-set
// dependent planes of examples/gestalt2d (PrevIterFeature: gaussian_filter(u, sigma) EVERY
// iteration), where eight gauss1d launches were half of the C2 step.
constexpr int G2_THREADS = 512;
constexpr int G2_WARPS = G2_THREADS / 32;

__device__ __forceinline__ double gauss_line(const double* __restrict__ line, int stride, int c,
                                             int n, const double* __restrict__ w, int r) {
    double tmp = line[c * stride] * w[r];
    if (c - r >= 0 && c + r < n) {
        for (int jj = -r; jj < 0; ++jj)
            tmp = tmp + (line[(c + jj) * stride] + line[(c - jj) * stride]) * w[r + jj];
    } else if (r <= n) {
        // one reflection at most (d c b a | a b c d | d c b a): no integer modulo on this path,
        // which every warp of the axis-1 pass takes (its lanes span interior and edge columns)
        for (int jj = -r; jj < 0; ++jj) {
            int ia = c + jj, ib = c - jj;
            ia = ia < 0 ? -ia - 1 : ia;
            ib = ib >= n ? 2 * n - 1 - ib : ib;
            tmp = tmp + (line[ia * stride] + line[ib * stride]) * w[r + jj];
        }
    } else {
        for (int jj = -r; jj < 0; ++jj)
            tmp = tmp + (line[reflect_index(c + jj, n) * stride] +
                         line[reflect_index(c - jj, n) * stride]) * w[r + jj];
    }
    return tmp;
}

// The item sits in shared memory once; every warp takes whole output rows: it first forms the
// row of the axis-0 pass (column taps: consecutive lanes read consecutive words) in a private
// row buffer, then runs the axis-1 pass along that buffer and stores the row -- no block
// barrier after the load, two CTAs per SM.
__global__ void __launch_bounds__(G2_THREADS)
gauss2d_smem_kernel(const double* __restrict__ in, double* __restrict__ out,
                    const double* __restrict__ w0, int r0, const double* __restrict__ w1, int r1,
                    int n1, int n2) {
    extern __shared__ double g2[];
    const int voxels = n1 * n2;
    double* a = g2;                                   // the item
    double* rows = g2 + voxels;                       // G2_WARPS row buffers of n2 doubles
    double* sw0 = rows + G2_WARPS * n2;               // weights, centre at [r]
    double* sw1 = sw0 + (r0 < 0 ? 0 : 2 * r0 + 1);
    const double* src = in + (i64)blockIdx.x * voxels;
    double* dst = out + (i64)blockIdx.x * voxels;
    for (int t = threadIdx.x; t < voxels; t += G2_THREADS) a[t] = __ldg(src + t);
    for (int t = threadIdx.x; t < 2 * r0 + 1; t += G2_THREADS) sw0[t] = __ldg(w0 + t);
    for (int t = threadIdx.x; t < 2 * r1 + 1; t += G2_THREADS) sw1[t] = __ldg(w1 + t);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* row = rows + warp * n2;
    // Four outputs per lane and step (columns lane, lane+32, lane+64, lane+96 of the warp's
    // row): four independent accumulation chains hide the float64 add latency of each.
    for (int i = warp; i < n1; i += G2_WARPS) {
        for (int j0 = lane; j0 < n2; j0 += 128) {
            int jq[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) jq[q] = min(j0 + 32 * q, n2 - 1);
            double t[4];
            if (r0 < 0) {                 // scipy skips sigma <= 1e-15
#pragma unroll
                for (int q = 0; q < 4; ++q) t[q] = a[i * n2 + jq[q]];
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) t[q] = a[i * n2 + jq[q]] * sw0[r0];
                const bool interior = i - r0 >= 0 && i + r0 < n1;
                for (int jj = -r0; jj < 0; ++jj) {
                    int ia = i + jj, ib = i - jj;
                    if (!interior) {
                        if (r0 <= n1) { ia = ia < 0 ? -ia - 1 : ia; ib = ib >= n1 ? 2 * n1 - 1 - ib : ib; }
                        else { ia = reflect_index(ia, n1); ib = reflect_index(ib, n1); }
                    }
                    const double wv = sw0[r0 + jj];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        t[q] = t[q] + (a[ia * n2 + jq[q]] + a[ib * n2 + jq[q]]) * wv;
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (j0 + 32 * q < n2) row[j0 + 32 * q] = t[q];
        }
        __syncwarp();
        for (int j0 = lane; j0 < n2; j0 += 128) {
            int jq[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) jq[q] = min(j0 + 32 * q, n2 - 1);
            double t[4];
            if (r1 < 0) {
#pragma unroll
                for (int q = 0; q < 4; ++q) t[q] = row[jq[q]];
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) t[q] = row[jq[q]] * sw1[r1];
                for (int jj = -r1; jj < 0; ++jj) {
                    const double wv = sw1[r1 + jj];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        int ia = jq[q] + jj, ib = jq[q] - jj;
                        if (r1 <= n2) { ia = ia < 0 ? -ia - 1 : ia; ib = ib >= n2 ? 2 * n2 - 1 - ib : ib; }
                        else { ia = reflect_index(ia, n2); ib = reflect_index(ib, n2); }
                        t[q] = t[q] + (row[ia] + row[ib]) * wv;
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (j0 + 32 * q < n2) dst[i * n2 + j0 + 32 * q] = t[q];
        }
        __syncwarp();
    }
}

// numpy.gradient along one axis (edge_order=1): interior (f[+1]-f[-1])/(2 dx), ends (f1-f0)/dx
__device__ __forceinline__ double np_grad_axis(const double* __restrict__ A, i64 l, i64 st, int c,
                                               int n, double dx) {
    if (n == 1) return 0.0;   // numpy raises for extent-1 axes; never reached from the host API
    if (c == 0) return (__ldg(A + l + st) - __ldg(A + l)) / dx;
    if (c == n - 1) return (__ldg(A + l) - __ldg(A + l - st)) / dx;
    return (__ldg(A + l + st) - __ldg(A + l - st)) / (2. * dx);
}

template <int NDIM>
__global__ void __launch_bounds__(256)
np_gradient_mag_kernel(const double* __restrict__ A, double* __restrict__ out, int n0, int n1,
                       int n2, i64 rows, double dx0, double dx1, double dx2) {
    const int k = blockIdx.z * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (k >= n2 || j >= n1) return;
    const i64 l = ((i64)blockIdx.x * n1 + j) * n2 + k;
    double acc = 0.0;   // reduce(lambda a, b: a + b**2, gradients, zeros) ** 0.5  (image.py:59-61)
    if (NDIM >= 3) {
        const double g = np_grad_axis(A, l, (i64)n1 * n2, (int)(blockIdx.x % (unsigned)n0), n0, dx0);
        acc = acc + g * g;
    }
    if (NDIM >= 2) {
        const double g = np_grad_axis(A, l, (i64)n2, j, n1, dx1);
        acc = acc + g * g;
    }
    const double g = np_grad_axis(A, l, (i64)1, k, n2, dx2);
    acc = acc + g * g;
    out[l] = sqrt(acc);
}

// ---------------------------------------------------------------------------------------------
// deterministic per-item sums over a dense array: stage 1 (slices x items), stage 2 (one warp)
// ---------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;

// mode 0: sum x ; mode 1: sum (x - mean[item])^2
__global__ void __launch_bounds__(RS_THREADS)
dense_sum_partial_kernel(const double* __restrict__ x, i64 voxels, int slices, int mode,
                         const double* __restrict__ mean, double* __restrict__ partial) {
    __shared__ double red[32];
    const int item = blockIdx.y, s = blockIdx.x;
    const i64 per = (voxels + slices - 1) / slices;
    const i64 lo = min(voxels, (i64)s * per), hi = min(voxels, lo + per);
    const double* p = x + (i64)item * voxels;
    const double mu = mode ? mean[item] : 0.0;
    double acc = 0.0;
    for (i64 i = lo + threadIdx.x; i < hi; i += RS_THREADS) {
        const double v = __ldg(p + i);
        if (mode) { const double d = v - mu; acc += d * d; } else acc += v;
    }
    const double tot = block_sum<double>(acc, red);
    if (threadIdx.x == 0) partial[(i64)item * slices + s] = tot;
}

// final: out[item] = f(sum partial) ; op 0: sum/count  op 1: sqrt(sum/count)
__global__ void sum_final_kernel(const double* __restrict__ partial, int slices, i64 batch,
                                 double count, int op, double* __restrict__ out) {
    const i64 item = (i64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= batch) return;
    const int lane = threadIdx.x & 31;
    double acc = 0.0;
    for (int s = lane; s < slices; s += 32) acc += partial[item * slices + s];
    acc = warp_sum(acc);
    if (lane == 0) out[item] = op == 0 ? acc / count : sqrt(acc / count);
}

__global__ void __launch_bounds__(256)
normalize_kernel(const double* __restrict__ img, double* __restrict__ out, i64 voxels,
                 const double* __restrict__ mean, const double* __restrict__ stdv) {
    const int item = blockIdx.y;
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= voxels) return;
    const i64 l = (i64)item * voxels + i;
    out[l] = (img[l] - mean[item]) / stdv[item];   // model.py:329
}

// ---------------------------------------------------------------------------------------------
// {u > 0} statistics as exact integers: stats[item][0]=count, [1..3]=sum idx_a, [4..6]=sum idx_a^2
// (padded axes), [7] = #(u == 0).  atomicAdd on integers is order-independent => deterministic.
// ---------------------------------------------------------------------------------------------
constexpr int GS_ROWS_PER_CTA = 32;

__global__ void __launch_bounds__(256)
global_stats_kernel(const double* __restrict__ u, int n0, int n1, int n2, u64* __restrict__ stats) {
    __shared__ u64 red[32];
    const int item = blockIdx.y;
    const i64 rows_per_item = (i64)n0 * n1;
    const i64 row0 = (i64)blockIdx.x * GS_ROWS_PER_CTA;
    const int bx = blockDim.x, by = blockDim.y;
    u64 cnt = 0, s0 = 0, s1 = 0, s2 = 0, q0 = 0, q1 = 0, q2 = 0, zeros = 0;
    for (int rr = threadIdx.y; rr < GS_ROWS_PER_CTA; rr += by) {
        const i64 row = row0 + rr;
        if (row >= rows_per_item) break;
        const u64 i = (u64)(row / n1), j = (u64)(row % n1);
        const double* p = u + ((i64)item * rows_per_item + row) * n2;
        for (int k = threadIdx.x; k < n2; k += bx) {
            const double v = __ldg(p + k);
            if (v > 0.0) {
                cnt += 1; s0 += i; s1 += j; s2 += (u64)k;
                q0 += i * i; q1 += j * j; q2 += (u64)k * (u64)k;
            } else if (v == 0.0) {
                zeros += 1;
            }
        }
    }
    u64 vals[8] = {cnt, s0, s1, s2, q0, q1, q2, zeros};
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const u64 t = block_sum<u64>(vals[q], red);
        if (threadIdx.x == 0 && threadIdx.y == 0 && t) atomicAdd(stats + (i64)item * 8 + q, t);
    }
}

// ---------------------------------------------------------------------------------------------
// Moments of order > 2 (shape.py:214-230 accepts any order >= 1).  The moment of order k along
// axis a is sum_{u>0} (i_a dx_a - c_a)^k * prod(dx) / measure: the summand depends on the voxel only
// through its index along that axis, so the EXACT per-axis histograms n_a[i] = #{u > 0, index_a
// = i} (integer atomics: deterministic) carry everything; the moment is then a sum over at most
// n_a bins.  Orders 1 and 2 never come here (exact integer statistics, maintained incrementally).
// ---------------------------------------------------------------------------------------------
constexpr int AH_ROWS_PER_CTA = 64;

__global__ void __launch_bounds__(256)
axis_hist_kernel(const double* __restrict__ u, int n0, int n1, int n2,
                 unsigned long long* __restrict__ hist /* [item][n0 + n1 + n2] */) {
    extern __shared__ unsigned sh[];          // n0 + n1 + n2 counters of this CTA
    const int item = blockIdx.y;
    const int nh = n0 + n1 + n2;
    for (int t = threadIdx.x; t < nh; t += blockDim.x) sh[t] = 0u;
    __syncthreads();
    const i64 rows_per_item = (i64)n0 * n1;
    const i64 row0 = (i64)blockIdx.x * AH_ROWS_PER_CTA;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int rr = warp; rr < AH_ROWS_PER_CTA; rr += nwarps) {
        const i64 row = row0 + rr;
        if (row >= rows_per_item) break;
        const int i = (int)(row / n1), j = (int)(row % n1);
        const double* p = u + ((i64)item * rows_per_item + row) * n2;
        unsigned cnt = 0u;
        for (int k = lane; k < n2; k += 32)
            if (__ldg(p + k) > 0.0) { ++cnt; atomicAdd(&sh[n0 + n1 + k], 1u); }
        cnt = warp_sum(cnt);
        if (lane == 0 && cnt) { atomicAdd(&sh[i], cnt); atomicAdd(&sh[n0 + j], cnt); }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nh; t += blockDim.x)
        if (sh[t]) atomicAdd(hist + (i64)item * nh + t, (unsigned long long)sh[t]);
}

// one warp per (item, requested column): item_feat[item][col] = moment of `order` along `axis`
__global__ void __launch_bounds__(32)
hi_moment_kernel(const unsigned long long* __restrict__ hist, int n0, int n1, int n2, ItemGeom geo,
                 const u64* __restrict__ stats, const double* __restrict__ com, int ncols,
                 const int* __restrict__ cols, const int* __restrict__ axes,
                 const int* __restrict__ orders, int nfeat, double* __restrict__ item_feat) {
    const i64 item = blockIdx.x / ncols;
    const int q = (int)(blockIdx.x % ncols);
    const int ax = geo.pad + axes[q], order = orders[q];
    const int ext[3] = {n0, n1, n2};
    const int off = ax == 0 ? 0 : (ax == 1 ? n0 : n0 + n1);
    const unsigned long long* h = hist + item * (i64)(n0 + n1 + n2) + off;
    const double c = com[item * 3 + ax];
    double acc = 0.0;
    for (int i = threadIdx.x; i < ext[ax]; i += 32) {
        const unsigned long long n = h[i];
        if (n == 0) continue;
        const double x = (double)(i + (ax == 0 ? geo.origin0 : 0)) * geo.dx[ax] - c;   // mesh - com
        double pw = x;
        for (int e = 1; e < order; ++e) pw *= x;
        acc += pw * (double)n;
    }
    acc = warp_sum(acc);
    if (threadIdx.x == 0) {
        const double measure = (double)stats[item * 8] * geo.pdx;
        item_feat[item * nfeat + cols[q]] = acc * geo.pdx / measure;
    }
}

// ---------------------------------------------------------------------------------------------
// band sums of P image planes, per item.  planes: (P, batch*voxels).  Deterministic two-stage:
// slice s of item i covers band positions [off_i + s*L, off_i + (s+1)*L), L = ceil(n_i/slices).
// mode 0: sum x -> partial ; mode 1: sum (x - mean[item][p])^2
// ---------------------------------------------------------------------------------------------
constexpr int MAX_PLANES = 16;

__global__ void __launch_bounds__(RS_THREADS)
band_sum_partial_kernel(const double* __restrict__ planes, i64 plane_stride, int nplanes,
                        const i64* __restrict__ band_idx, const i64* __restrict__ item_offsets,
                        int slices, int mode, const double* __restrict__ mean /* [item][P] */,
                        double* __restrict__ partial /* [item][P][slices] */) {
    __shared__ double red[32];
    const int item = blockIdx.y, s = blockIdx.x;
    const i64 off = item_offsets[item], n = item_offsets[item + 1] - off;
    const i64 per = (n + slices - 1) / slices;
    const i64 lo = min(n, (i64)s * per), hi = min(n, lo + per);
    double acc[MAX_PLANES];
#pragma unroll
    for (int p = 0; p < MAX_PLANES; ++p) acc[p] = 0.0;
    for (i64 b = lo + threadIdx.x; b < hi; b += RS_THREADS) {
        const i64 g = __ldg(band_idx + off + b);
#pragma unroll
        for (int p = 0; p < MAX_PLANES; ++p) {
            if (p < nplanes) {
                const double v = __ldg(planes + (i64)p * plane_stride + g);
                if (mode) { const double d = v - mean[(i64)item * nplanes + p]; acc[p] += d * d; }
                else acc[p] += v;
            }
        }
    }
#pragma unroll
    for (int p = 0; p < MAX_PLANES; ++p) {
        if (p < nplanes) {
            const double t = block_sum<double>(acc[p], red);
            if (threadIdx.x == 0) partial[((i64)item * nplanes + p) * slices + s] = t;
        }
    }
}

// out[item][p] = op(sum_s partial / n_item) ; op 0 mean, op 1 sqrt (population std, ddof=0),
// op 2 the raw sum (slab decomposition: the ranks' sums are all-reduced first)
__global__ void band_sum_final_kernel(const double* __restrict__ partial, int nplanes, int slices,
                                      const i64* __restrict__ item_offsets, i64 batch, int op,
                                      double* __restrict__ out) {
    const i64 w = (i64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= batch * nplanes) return;
    const i64 item = w / nplanes;
    const int lane = threadIdx.x & 31;
    double acc = 0.0;
    for (int s = lane; s < slices; s += 32) acc += partial[w * slices + s];
    acc = warp_sum(acc);
    if (lane == 0) {
        const double n = (double)(item_offsets[item + 1] - item_offsets[item]);
        out[w] = op == 2 ? acc : (op == 0 ? acc / n : sqrt(acc / n));
    }
}

// ---------------------------------------------------------------------------------------------
// marching squares (skimage find_contours level 0, fully_connected='low') total segment length,
// scaled like lsml: row differences * dx[1], column differences * dx[0] (shape.py:80).
// Closed contours only contribute their segments; the closing chord of OPEN contours
// (shape.py:79) is added by contour_chord_kernel below.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double ms_frac(double from, double to) {
    return to == from ? 0.0 : (0.0 - from) / (to - from);
}
__device__ __forceinline__ double seg_len(double r0, double c0, double r1, double c1, double sr,
                                          double sc) {
    const double dr = (r1 - r0) * sr, dc = (c1 - c0) * sc;
    return sqrt(dr * dr + dc * dc);
}

__global__ void __launch_bounds__(256)
contour_length_kernel(const double* __restrict__ u, int m, int n, double sr, double sc,
                      int cells_per_cta, double* __restrict__ partial) {
    __shared__ double red[32];
    const int item = blockIdx.y;
    const double* p = u + (i64)item * m * n;
    const i64 ncells = (i64)(m - 1) * (n - 1);
    const i64 lo = (i64)blockIdx.x * cells_per_cta;
    const i64 hi = min(ncells, lo + cells_per_cta);
    double acc = 0.0;
    for (i64 c = lo + threadIdx.x; c < hi; c += blockDim.x) {
        const int r0 = (int)(c / (n - 1)), c0 = (int)(c % (n - 1));
        const double ul = __ldg(p + (i64)r0 * n + c0), ur = __ldg(p + (i64)r0 * n + c0 + 1);
        const double ll = __ldg(p + (i64)(r0 + 1) * n + c0), lr = __ldg(p + (i64)(r0 + 1) * n + c0 + 1);
        if (ul != ul || ur != ur || ll != ll || lr != lr) continue;
        const int sq = (ul > 0) + 2 * (ur > 0) + 4 * (ll > 0) + 8 * (lr > 0);
        if (sq == 0 || sq == 15) continue;
        const double tr = r0, tc = c0 + ms_frac(ul, ur);          // top
        const double br = r0 + 1, bc = c0 + ms_frac(ll, lr);      // bottom
        const double lr_ = r0 + ms_frac(ul, ll), lc = c0;         // left
        const double rr = r0 + ms_frac(ur, lr), rc = c0 + 1;      // right
        switch (sq) {
            case 1: case 14: acc += seg_len(tr, tc, lr_, lc, sr, sc); break;
            case 2: case 13: acc += seg_len(rr, rc, tr, tc, sr, sc); break;
            case 3: case 12: acc += seg_len(rr, rc, lr_, lc, sr, sc); break;
            case 4: case 11: acc += seg_len(lr_, lc, br, bc, sr, sc); break;
            case 5: case 10: acc += seg_len(tr, tc, br, bc, sr, sc); break;
            case 7: case 8: acc += seg_len(rr, rc, br, bc, sr, sc); break;
            case 6: acc += seg_len(rr, rc, tr, tc, sr, sc); acc += seg_len(lr_, lc, br, bc, sr, sc); break;
            case 9: acc += seg_len(tr, tc, lr_, lc, sr, sc); acc += seg_len(br, bc, rr, rc, sr, sc); break;
        }
    }
    const double t = block_sum<double>(acc, red);
    if (threadIdx.x == 0) partial[(i64)item * gridDim.x + blockIdx.x] = t;
}

// [host-test:begin] (tests/test_feature_kernels_host.py compiles this region with g++, after the
// marked region of band_features.cuh: the two kernels below have no block-level cooperation, so
// one host "thread" can run them as they are)
__global__ void item_features_kernel(FeatProgram prog, ItemGeom geo, i64 batch, int nplanes,
                                     const u64* __restrict__ stats,
                                     const double* __restrict__ band_mean,
                                     const double* __restrict__ band_std,
                                     const double* __restrict__ boundary,
                                     double* __restrict__ item_feat /* [item][nfeat] */,
                                     double* __restrict__ com /* [item][3] padded axes */) {
    const i64 item = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= batch) return;
    const u64* st = stats + item * 8;
    for (int ax = 0; ax < 3; ++ax)
        com[item * 3 + ax] = ax >= geo.pad ? moment_from_stats(st, ax, 1, geo.dx[ax], geo.pdx) : 0.0;
    const double size = (double)st[0] * geo.pdx;
    const double bsz = boundary ? boundary[item] : 0.0;
    const double pi = 3.141592653589793;
    for (int f = 0; f < prog.nfeat; ++f) {
        double v = 0.0;
        switch (prog.kind[f]) {
            case FK_BAND_MEAN: v = band_mean[item * nplanes + prog.a[f]]; break;
            case FK_BAND_STD: v = band_std[item * nplanes + prog.a[f]]; break;
            case FK_SIZE: v = size; break;
            case FK_BOUNDARY: v = bsz; break;
            case FK_ISOPERIMETRIC:
                v = geo.ndim == 2 ? 4 * pi * size / (bsz * bsz)                 // shape.py:136
                                  : 36 * pi * (size * size) / (bsz * bsz * bsz); // shape.py:151
                break;
            case FK_MOMENT: {
                const int ax = geo.pad + prog.a[f];
                v = moment_from_stats(st, ax, prog.b[f], geo.dx[ax], geo.pdx);
                break;
            }
            default: break;
        }
        item_feat[item * prog.nfeat + f] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// (B, F) feature matrix in band order.  One thread per (band voxel, feature) pair with the
// feature index fastest => fully coalesced 8-byte stores of the row-major matrix.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
assemble_kernel(BandFeatures src, const i64* __restrict__ n_band, double* __restrict__ X) {
    const i64 nb = *n_band;
    const i64 total = nb * src.prog.nfeat;
    for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (i64)gridDim.x * blockDim.x) {
        const i64 b = t / src.prog.nfeat;
        const int f = (int)(t - b * src.prog.nfeat);
        X[t] = band_feature(src, voxel_ref(src, __ldg(src.band_idx + b)), f);
    }
}
// [host-test:end]

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
static void rows_launch(const Geom& g, dim3& grid, dim3& block, i64& rows) {
    int bx = 32;
    while (bx < 256 && bx < g.n[2]) bx <<= 1;
    int by = 256 / bx;
    if (by > g.n[1]) { by = 1; while (by < g.n[1]) by <<= 1; }
    rows = g.batch * g.n[0] * g.n[1];
    block = dim3(bx, by, 1);
    grid = dim3((unsigned)(g.batch * g.n[0]), (unsigned)((g.n[1] + by - 1) / by),
                (unsigned)((g.n[2] + bx - 1) / bx));
}

template <int R>
static void gauss_march_launch(const double* in, double* out, const double* w, int antisym, int n,
                               i64 st, i64 inner, i64 outer, i64 outer_stride, int mode,
                               cudaStream_t s) {
    const dim3 grid((unsigned)outer, (unsigned)((inner + GM_THREADS - 1) / GM_THREADS));
    gauss1d_march_kernel<R><<<grid, GM_THREADS, 0, s>>>(in, out, w, antisym, n, st, inner,
                                                        outer_stride, mode);
}

int gauss1d(const Geom& g, const double* in, double* out, const double* w, int radius,
            int antisym, int axis_padded, int mode, cudaStream_t s) {
    // strided axes: the register march (in-place filtering is not possible there: a column's
    // outputs overwrite inputs the ring has not loaded yet only when in == out)
    static const bool no_march = getenv("LSML_GAUSS_NO_MARCH") != nullptr;      // A/B switch (probes)
    if (!no_march && axis_padded < 2 && in != out && radius >= 1 && radius <= 16 && g.n[axis_padded] > 1) {
        const i64 plane = (i64)g.n[1] * g.n[2];
        const int n = g.n[axis_padded];
        const i64 st = axis_padded == 0 ? plane : g.n[2];
        const i64 inner = axis_padded == 0 ? plane : g.n[2];
        const i64 outer = axis_padded == 0 ? g.batch : g.batch * g.n[0];
        const i64 ostride = axis_padded == 0 ? g.voxels : plane;
        if (inner >= 32 && outer < ((i64)1 << 31) && (inner + GM_THREADS - 1) / GM_THREADS <= 65535) {
            switch (radius) {
#define LSML_GM(R) case R: gauss_march_launch<R>(in, out, w, antisym, n, st, inner, outer, ostride, mode, s); break;
                LSML_GM(1) LSML_GM(2) LSML_GM(3) LSML_GM(4) LSML_GM(5) LSML_GM(6) LSML_GM(7) LSML_GM(8)
                LSML_GM(9) LSML_GM(10) LSML_GM(11) LSML_GM(12) LSML_GM(13) LSML_GM(14) LSML_GM(15) LSML_GM(16)
#undef LSML_GM
            }
            LSML_LAUNCH_CHECK("gauss1d_march_kernel");
            return 0;
        }
    }
    dim3 grid, block; i64 rows;
    rows_launch(g, grid, block, rows);
    gauss1d_kernel<<<grid, block, 0, s>>>(in, out, w, radius, antisym, g.n[0], g.n[1], g.n[2],
                                          axis_padded, rows, mode);
    LSML_LAUNCH_CHECK("gauss1d_kernel");
    return 0;
}

// gaussian_filter(in, (sigma0, sigma1)) for 2-D items that fit in shared memory twice; radius < 0
// skips an axis.  *done = false when the items are too large (the caller chains gauss1d).
int gauss2d_small(const Geom& g, const double* in, double* out, const double* w0, int r0,
                  const double* w1, int r1, bool* done, cudaStream_t s) {
    *done = false;
    if (g.ndim != 2) return 0;
    const int q0 = r0 < 0 ? 0 : 2 * r0 + 1, q1 = r1 < 0 ? 0 : 2 * r1 + 1;
    const size_t smem = ((size_t)g.voxels + (size_t)G2_WARPS * g.n[2] + q0 + q1 + 2) * sizeof(double);
    if (smem > 220 * 1024 || g.batch > 0x7fffffff) return 0;
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        LSML_CUDA(cudaFuncSetAttribute(gauss2d_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    gauss2d_smem_kernel<<<(unsigned)g.batch, G2_THREADS, smem, s>>>(in, out, w0, r0, w1, r1, g.n[1], g.n[2]);
    LSML_LAUNCH_CHECK("gauss2d_smem_kernel");
    *done = true;
    return 0;
}

int np_gradient_mag(const Geom& g, const double* in, double* out, cudaStream_t s) {
    dim3 grid, block; i64 rows;
    rows_launch(g, grid, block, rows);
    if (g.ndim == 3)
        np_gradient_mag_kernel<3><<<grid, block, 0, s>>>(in, out, g.n[0], g.n[1], g.n[2], rows, g.dx[0], g.dx[1], g.dx[2]);
    else if (g.ndim == 2)
        np_gradient_mag_kernel<2><<<grid, block, 0, s>>>(in, out, g.n[0], g.n[1], g.n[2], rows, g.dx[0], g.dx[1], g.dx[2]);
    else
        np_gradient_mag_kernel<1><<<grid, block, 0, s>>>(in, out, g.n[0], g.n[1], g.n[2], rows, g.dx[0], g.dx[1], g.dx[2]);
    LSML_LAUNCH_CHECK("np_gradient_mag_kernel");
    return 0;
}

static int dense_slices(i64 voxels) {
    i64 s = (voxels + 16383) / 16384;
    return (int)(s < 1 ? 1 : (s > 1024 ? 1024 : s));
}

i64 normalize_workspace_bytes(i64 voxels, i64 batch) {
    return (i64)sizeof(double) * (batch * dense_slices(voxels) + 2 * batch);
}

int normalize_images(i64 voxels, i64 batch, const double* img, double* out, void* workspace,
                     cudaStream_t s) {
    const int slices = dense_slices(voxels);
    double* partial = static_cast<double*>(workspace);
    double* mean = partial + batch * slices;
    double* stdv = mean + batch;
    const dim3 grid(slices, (unsigned)batch);
    const int fin_blocks = (int)((batch + 7) / 8);
    dense_sum_partial_kernel<<<grid, RS_THREADS, 0, s>>>(img, voxels, slices, 0, nullptr, partial);
    LSML_LAUNCH_CHECK("dense_sum_partial_kernel");
    sum_final_kernel<<<fin_blocks, 256, 0, s>>>(partial, slices, batch, (double)voxels, 0, mean);
    LSML_LAUNCH_CHECK("sum_final_kernel");
    dense_sum_partial_kernel<<<grid, RS_THREADS, 0, s>>>(img, voxels, slices, 1, mean, partial);
    LSML_LAUNCH_CHECK("dense_sum_partial_kernel");
    sum_final_kernel<<<fin_blocks, 256, 0, s>>>(partial, slices, batch, (double)voxels, 1, stdv);
    LSML_LAUNCH_CHECK("sum_final_kernel");
    const dim3 ngrid((unsigned)((voxels + 255) / 256), (unsigned)batch);
    normalize_kernel<<<ngrid, 256, 0, s>>>(img, out, voxels, mean, stdv);
    LSML_LAUNCH_CHECK("normalize_kernel");
    return 0;
}

int global_stats(const Geom& g, const double* u, u64* stats, cudaStream_t s) {
    LSML_CUDA(cudaMemsetAsync(stats, 0, sizeof(u64) * 8 * g.batch, s));
    int bx = 32;
    while (bx < 256 && bx < g.n[2]) bx <<= 1;
    const dim3 block(bx, 256 / bx);
    const i64 rows_per_item = (i64)g.n[0] * g.n[1];
    const dim3 grid((unsigned)((rows_per_item + GS_ROWS_PER_CTA - 1) / GS_ROWS_PER_CTA), (unsigned)g.batch);
    global_stats_kernel<<<grid, block, 0, s>>>(u, g.n[0], g.n[1], g.n[2], stats);
    LSML_LAUNCH_CHECK("global_stats_kernel");
    return 0;
}

int band_slices(i64 voxels) {
    i64 s = (voxels + 32767) / 32768;
    return (int)(s < 1 ? 1 : (s > 512 ? 512 : s));
}

i64 band_stats_workspace_bytes(i64 voxels, i64 batch, int nplanes) {
    return (i64)sizeof(double) * batch * nplanes * band_slices(voxels);
}

// mean and population std of every plane over each item's band
int band_stats(i64 voxels, i64 batch, const double* planes, i64 plane_stride, int nplanes,
               const i64* band_idx, const i64* item_offsets, double* mean, double* stdv,
               void* workspace, cudaStream_t s) {
    if (nplanes > MAX_PLANES) { set_error("at most %d image planes", MAX_PLANES); return 1; }
    if (nplanes == 0) return 0;
    const int slices = band_slices(voxels);
    double* partial = static_cast<double*>(workspace);
    const dim3 grid(slices, (unsigned)batch);
    const int fin_blocks = (int)((batch * nplanes + 7) / 8);
    band_sum_partial_kernel<<<grid, RS_THREADS, 0, s>>>(planes, plane_stride, nplanes, band_idx,
                                                        item_offsets, slices, 0, nullptr, partial);
    LSML_LAUNCH_CHECK("band_sum_partial_kernel");
    band_sum_final_kernel<<<fin_blocks, 256, 0, s>>>(partial, nplanes, slices, item_offsets, batch, 0, mean);
    LSML_LAUNCH_CHECK("band_sum_final_kernel");
    band_sum_partial_kernel<<<grid, RS_THREADS, 0, s>>>(planes, plane_stride, nplanes, band_idx,
                                                        item_offsets, slices, 1, mean, partial);
    LSML_LAUNCH_CHECK("band_sum_partial_kernel");
    band_sum_final_kernel<<<fin_blocks, 256, 0, s>>>(partial, nplanes, slices, item_offsets, batch, 1, stdv);
    LSML_LAUNCH_CHECK("band_sum_final_kernel");
    return 0;
}

// raw per-item band sums of every plane: mode 0 sum x, mode 1 sum (x - mean)^2 with the given mean
int band_sums(i64 voxels, i64 batch, const double* planes, i64 plane_stride, int nplanes,
              const i64* band_idx, const i64* item_offsets, int mode, const double* mean,
              double* sums, void* workspace, cudaStream_t s) {
    if (nplanes > MAX_PLANES) { set_error("at most %d image planes", MAX_PLANES); return 1; }
    if (nplanes == 0) return 0;
    const int slices = band_slices(voxels);
    double* partial = static_cast<double*>(workspace);
    const dim3 grid(slices, (unsigned)batch);
    const int fin_blocks = (int)((batch * nplanes + 7) / 8);
    band_sum_partial_kernel<<<grid, RS_THREADS, 0, s>>>(planes, plane_stride, nplanes, band_idx,
                                                        item_offsets, slices, mode, mean, partial);
    LSML_LAUNCH_CHECK("band_sum_partial_kernel");
    band_sum_final_kernel<<<fin_blocks, 256, 0, s>>>(partial, nplanes, slices, item_offsets, batch, 2, sums);
    LSML_LAUNCH_CHECK("band_sum_final_kernel");
    return 0;
}

constexpr int CL_CELLS_PER_CTA = 8192;

i64 contour_workspace_bytes(int m, int n, i64 batch) {
    const i64 ncells = (i64)(m - 1) * (n - 1);
    const i64 ctas = (ncells + CL_CELLS_PER_CTA - 1) / CL_CELLS_PER_CTA;
    return (i64)sizeof(double) * batch * (ctas < 1 ? 1 : ctas);
}

// 2-D only: boundary[item] = total marching-squares length (closed contours)
int contour_length(const Geom& g, const double* u, double* boundary, void* workspace,
                   cudaStream_t s) {
    if (g.ndim != 2) { set_error("contour_length: 2-D grids only"); return 1; }
    const int m = g.n[1], n = g.n[2];
    const i64 ncells = (i64)(m - 1) * (n - 1);
    i64 ctas = (ncells + CL_CELLS_PER_CTA - 1) / CL_CELLS_PER_CTA;
    if (ctas < 1) ctas = 1;
    double* partial = static_cast<double*>(workspace);
    // shape.py:80: (row, col) * dx[::-1]  => rows scale with dx[1], columns with dx[0]
    contour_length_kernel<<<dim3((unsigned)ctas, (unsigned)g.batch), 256, 0, s>>>(
        u, m, n, g.dx[2], g.dx[1], CL_CELLS_PER_CTA, partial);
    LSML_LAUNCH_CHECK("contour_length_kernel");
    sum_final_kernel<<<(int)((g.batch + 7) / 8), 256, 0, s>>>(partial, (int)ctas, g.batch, 1.0, 0, boundary);
    LSML_LAUNCH_CHECK("sum_final_kernel");
    return 0;
}

int item_features(const FeatProgram& prog, const Geom& g, int nplanes, const u64* stats,
                  const double* band_mean, const double* band_std, const double* boundary,
                  double* item_feat, double* com, cudaStream_t s) {
    ItemGeom geo;
    geo.ndim = g.ndim; geo.pad = 3 - g.ndim;
    double pdx = 1.0;
    for (int a = 0; a < 3; ++a) geo.dx[a] = g.dx[a];
    for (int a = geo.pad; a < 3; ++a) pdx = (a == geo.pad) ? g.dx[a] : pdx * g.dx[a];
    geo.pdx = pdx; geo.origin0 = 0; geo.n0 = g.n[0];
    item_features_kernel<<<(unsigned)((g.batch + 127) / 128), 128, 0, s>>>(
        prog, geo, g.batch, nplanes, stats, band_mean, band_std, boundary, item_feat, com);
    LSML_LAUNCH_CHECK("item_features_kernel");
    return 0;
}

BandFeatures make_band_features(const FeatProgram& prog, const Geom& g, const i64* band_idx,
                                const double* planes, i64 plane_stride, const double* item_feat,
                                const double* com, int origin0) {
    BandFeatures src;
    src.prog = prog;
    src.geo.ndim = g.ndim; src.geo.pad = 3 - g.ndim;
    for (int a = 0; a < 3; ++a) src.geo.dx[a] = g.dx[a];
    src.geo.pdx = 1.0; src.geo.origin0 = origin0; src.geo.n0 = g.n[0];
    src.n1 = g.n[1]; src.n2 = g.n[2]; src.voxels = g.voxels;
    src.band_idx = band_idx; src.planes = planes; src.plane_stride = plane_stride;
    src.item_feat = item_feat; src.com = com;
    return src;
}

int assemble_features(const FeatProgram& prog, const Geom& g, const i64* band_idx,
                      const i64* n_band, const double* planes, i64 plane_stride,
                      const double* item_feat, const double* com, double* X, int origin0,
                      cudaStream_t s) {
    const BandFeatures src = make_band_features(prog, g, band_idx, planes, plane_stride, item_feat,
                                                com, origin0);
    assemble_kernel<<<LSML_SMS * 8, 256, 0, s>>>(src, n_band, X);
    LSML_LAUNCH_CHECK("assemble_kernel");
    return 0;
}

}  // namespace lsml

namespace lsml {

int fill_program(FeatProgram& p, int nfeat, const int* kind, const int* a, const int* b) {
    if (nfeat < 0 || nfeat > 64) { set_error("feature program: 0..64 columns (got %d)", nfeat); return 1; }
    p.nfeat = nfeat;
    for (int f = 0; f < nfeat; ++f) {
        p.kind[f] = kind[f]; p.a[f] = a[f]; p.b[f] = b[f];
        if (kind[f] == FK_MOMENT && (b[f] < 1 || b[f] > 64)) {
            set_error("Moments order should be greater than or equal to 1 (and <= 64; got %d)", b[f]);
            return 1;
        }
        if (kind[f] == FK_COM_RAY && ((b[f] >> 16) < 1 || (b[f] & 0xffff) >= 2 * (b[f] >> 16))) {
            set_error("COMRaySample: sample %d of n_samples %d", b[f] & 0xffff, b[f] >> 16);
            return 1;
        }
    }
    return 0;
}

int item_features_c(int nfeat, const int* kind, const int* a, const int* b, const Geom& g,
                    int nplanes, const u64* stats, const double* band_mean,
                    const double* band_std, const double* boundary, double* item_feat,
                    double* com, cudaStream_t s) {
    FeatProgram p;
    if (fill_program(p, nfeat, kind, a, b)) return 1;
    return item_features(p, g, nplanes, stats, band_mean, band_std, boundary, item_feat, com, s);
}

int assemble_features_c(int nfeat, const int* kind, const int* a, const int* b, const Geom& g,
                        const i64* band_idx, const i64* n_band, const double* planes,
                        i64 plane_stride, const double* item_feat, const double* com, double* X,
                        int origin0, cudaStream_t s) {
    FeatProgram p;
    if (fill_program(p, nfeat, kind, a, b)) return 1;
    for (int f = 0; f < nfeat; ++f) {
        if (kind[f] != FK_COM_RAY) continue;
        // the ray leaves the voxel's neighbourhood: the whole image must be local
        if (origin0 != 0) { set_error("COMRaySample is not available on slab-decomposed volumes"); return 1; }
        if (g.ndim < 2) { set_error("COMRaySample: `ndim` must be either 2 or 3"); return 1; }
        for (int ax = 3 - g.ndim; ax < 3; ++ax)
            if (g.n[ax] < 2) { set_error("COMRaySample: every axis needs at least 2 grid points"); return 1; }
    }
    return assemble_features(p, g, band_idx, n_band, planes, plane_stride, item_feat, com, X,
                             origin0, s);
}

i64 axis_moments_workspace_bytes(const Geom& g) {
    return (i64)sizeof(unsigned long long) * g.batch * (g.n[0] + g.n[1] + g.n[2]);
}

// Moments of order > 2 into their columns of item_feat (after item_features has filled com).
// cols / axes / orders: DEVICE int arrays of length ncols.
int axis_moments(const Geom& g, const double* u, const u64* stats, const double* com, int ncols,
                 const int* cols, const int* axes, const int* orders, int nfeat, double* item_feat,
                 void* workspace, cudaStream_t s) {
    if (ncols <= 0) return 0;
    if (g.batch > 65535) { set_error("axis_moments: at most 65535 items per call"); return 1; }
    const int nh = g.n[0] + g.n[1] + g.n[2];
    if ((size_t)nh * sizeof(unsigned) > 200 * 1024) { set_error("axis_moments: grid too large"); return 1; }
    unsigned long long* hist = static_cast<unsigned long long*>(workspace);
    LSML_CUDA(cudaMemsetAsync(hist, 0, sizeof(unsigned long long) * g.batch * nh, s));
    const i64 rows = (i64)g.n[0] * g.n[1];
    const unsigned ctas = (unsigned)((rows + AH_ROWS_PER_CTA - 1) / AH_ROWS_PER_CTA);
    const size_t smem = (size_t)nh * sizeof(unsigned);
    if (smem > 48 * 1024)
        LSML_CUDA(cudaFuncSetAttribute(axis_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
    axis_hist_kernel<<<dim3(ctas, (unsigned)g.batch), 256, smem, s>>>(u, g.n[0], g.n[1], g.n[2], hist);
    LSML_LAUNCH_CHECK("axis_hist_kernel");
    ItemGeom geo;
    geo.ndim = g.ndim; geo.pad = 3 - g.ndim;
    double pdx = 1.0;
    for (int a = 0; a < 3; ++a) geo.dx[a] = g.dx[a];
    for (int a = geo.pad; a < 3; ++a) pdx = (a == geo.pad) ? g.dx[a] : pdx * g.dx[a];
    geo.pdx = pdx; geo.origin0 = 0; geo.n0 = g.n[0];
    hi_moment_kernel<<<(unsigned)(g.batch * ncols), 32, 0, s>>>(hist, g.n[0], g.n[1], g.n[2], geo,
                                                                stats, com, ncols, cols, axes,
                                                                orders, nfeat, item_feat);
    LSML_LAUNCH_CHECK("hi_moment_kernel");
    return 0;
}

}  // namespace lsml
