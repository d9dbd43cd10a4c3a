// update.cu -- the banded level-set update (SURVEY.md section 8a rows a1 (banded) + a13).
//
//   gmag = gradient_magnitude_osher_sethian(u, velocity, mask, dx)     model.py:390
//   u[mask] += step * velocity[mask] * gmag[mask]                      model.py:394
//
// fused per band voxel: the upwind |Du| is evaluated from the PRE-update u (quirk 1 of
// SURVEY.md section 9), the product is (step*v)*gmag left to right (quirk 13).  Because band
// voxels are neighbours of each other the new values go to a banded buffer first
// (band_update_kernel) and are scattered into u by band_apply_kernel; the banded buffer is
// also the sparse per-iteration record from which `segment` rebuilds the iterates `us`.
//
// Algorithmic bytes per band voxel (f64): 8 (index) + 8 (v) + 8*(1 + H/B) (u incl. one-ring
// halo) + 8 (new value) + 8 + 8 (apply: read new value, write u) ~ 56-64 B.
#include "lsml_internal.h"

namespace lsml {

__device__ __forceinline__ void fwd_bwd_b(const double* __restrict__ A, i64 l, i64 st, int c,
                                          int n, double center, double& f, double& b) {
    if (n == 1) { f = 0.0; b = 0.0; return; }
    if (c == 0) { f = __ldg(A + l + st) - center; b = f; }
    else if (c == n - 1) { b = center - __ldg(A + l - st); f = b; }
    else { f = __ldg(A + l + st) - center; b = center - __ldg(A + l - st); }
}

__device__ __forceinline__ double os_acc(double b, double f, bool neg, double acc, bool first) {
    double tb, tf;
    if (neg) { tb = ref_max(b, 0.0); tf = ref_min(f, 0.0); }
    else     { tb = ref_min(b, 0.0); tf = ref_max(f, 0.0); }
    tb = tb * tb; tf = tf * tf;
    return first ? (tb + tf) : ((acc + tb) + tf);
}

struct BandGeom {
    int ndim, n0, n1, n2;
    i64 voxels;
    double dx0, dx1, dx2, r0, r1, r2;
    double inv_voxels;     // 1 / voxels (item of a flat index without a 64-bit division)
    int small;             // voxels < 2^32: the in-item index fits 32 bits
    int single;            // batch == 1
};

// flat index -> (item, i, j, k).  64-bit integer divisions cost ~100 instructions each on this
// path (four per voxel before); the item comes from one multiplication by 1/voxels with a +-1
// fix-up (exact: l < 2^53), the coordinates from 32-bit divisions when the item fits 32 bits.
__device__ __forceinline__ i64 decode_voxel(i64 l, const BandGeom& g, int& i, int& j, int& k) {
    i64 item = 0, loc = l;
    if (!g.single) {
        item = __double2ll_rz((double)l * g.inv_voxels);
        loc = l - item * g.voxels;
        if (loc < 0) { loc += g.voxels; --item; }
        else if (loc >= g.voxels) { loc -= g.voxels; ++item; }
    }
    if (g.small) {
        const unsigned lo = (unsigned)loc;
        const unsigned r = lo / (unsigned)g.n2;
        k = (int)(lo - r * (unsigned)g.n2);
        const unsigned ii = r / (unsigned)g.n1;
        j = (int)(r - ii * (unsigned)g.n1);
        i = (int)ii;
    } else {
        k = (int)(loc % g.n2);
        const i64 r = loc / g.n2;
        j = (int)(r % g.n1); i = (int)(r / g.n1);
    }
    return item;
}

__global__ void __launch_bounds__(256)
band_update_kernel(const double* __restrict__ u, const double* __restrict__ vel,
                   const i64* __restrict__ band_idx, const i64* __restrict__ n_band, BandGeom g,
                   double step, double* __restrict__ new_u, double* __restrict__ gmag_out) {
    const i64 nb = *n_band;
    for (i64 b = (i64)blockIdx.x * blockDim.x + threadIdx.x; b < nb;
         b += (i64)gridDim.x * blockDim.x) {
        const i64 l = __ldg(band_idx + b);
        int i, j, k;
        decode_voxel(l, g, i, j, k);
        const double c = __ldg(u + l), v = __ldg(vel + b);
        const bool neg = v < 0.0;
        double acc = 0.0, f, bk;
        bool first = true;
        if (g.ndim >= 3) {
            fwd_bwd_b(u, l, (i64)g.n1 * g.n2, i, g.n0, c, f, bk);
            f = div_dx(f, g.dx0, g.r0); bk = div_dx(bk, g.dx0, g.r0);
            acc = os_acc(bk, f, neg, acc, first); first = false;
        }
        if (g.ndim >= 2) {
            fwd_bwd_b(u, l, (i64)g.n2, j, g.n1, c, f, bk);
            f = div_dx(f, g.dx1, g.r1); bk = div_dx(bk, g.dx1, g.r1);
            acc = os_acc(bk, f, neg, acc, first); first = false;
        }
        fwd_bwd_b(u, l, (i64)1, k, g.n2, c, f, bk);
        f = div_dx(f, g.dx2, g.r2); bk = div_dx(bk, g.dx2, g.r2);
        acc = os_acc(bk, f, neg, acc, first);
        const double gm = sqrt(acc);
        if (gmag_out != nullptr) gmag_out[b] = gm;
        new_u[b] = c + (step * v) * gm;
    }
}

// Scatter of the new values.  u changes only here, so the exact {u > 0} statistics that Size /
// Moments / centre of mass need (features.cu:global_stats_kernel -- integer sums, hence
// order-independent) are kept up to date from the voxels that change sign class instead of a
// dense pass over the grid every iteration: stats[item*8 + {0 count, 1..3 sum idx, 4..6 sum
// idx^2, 7 zeros}].
// The per-thread deltas are reduced over the CTA before they touch the counters (same-address
// atomics serialise in L2: one atomic per flipped voxel cost 90 us at 256^3).  Band indices are
// ascending, so a CTA's contiguous chunk of the band belongs to one item except at item
// boundaries, where the partial sums are flushed early.
constexpr int APPLY_THREADS = 256;

__device__ __forceinline__ void flush_stats(u64 (&d)[8], i64 item, u64* __restrict__ stats) {
    __shared__ u64 red[32];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const u64 t = block_sum<u64>(d[q], red);
        if (threadIdx.x == 0 && t) atomicAdd(stats + item * 8 + q, t);
        d[q] = 0;
    }
}

__global__ void __launch_bounds__(APPLY_THREADS)
band_apply_kernel(double* __restrict__ u, const double* __restrict__ new_u,
                  const i64* __restrict__ band_idx, const i64* __restrict__ n_band, BandGeom g,
                  u64* __restrict__ stats) {
    const i64 nb = *n_band;
    if (stats == nullptr) {
        for (i64 b = (i64)blockIdx.x * blockDim.x + threadIdx.x; b < nb;
             b += (i64)gridDim.x * blockDim.x)
            u[__ldg(band_idx + b)] = __ldg(new_u + b);
        return;
    }
    // contiguous chunk of the band per CTA
    const i64 per = (nb + gridDim.x - 1) / gridDim.x;
    const i64 lo = min(nb, (i64)blockIdx.x * per), hi = min(nb, lo + per);
    if (lo >= hi) return;
    const i64 first_item = __ldg(band_idx + lo) / g.voxels;
    const i64 last_item = __ldg(band_idx + hi - 1) / g.voxels;
    for (i64 item = first_item; item <= last_item; ++item) {      // one pass in all but edge cases
        u64 d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const i64 ibase = item * g.voxels;
        for (i64 b = lo + threadIdx.x; b < hi; b += APPLY_THREADS) {
            const i64 l = __ldg(band_idx + b);
            const i64 loc = l - ibase;
            if (loc < 0 || loc >= g.voxels) continue;             // another item's voxel
            const double nv = __ldg(new_u + b), ov = u[l];
            const int pos = (nv > 0.0) - (ov > 0.0), zer = (nv == 0.0) - (ov == 0.0);
            if (pos) {
                int ci, cj, ck;
                decode_voxel(l, g, ci, cj, ck);
                const u64 i = (u64)ci, j = (u64)cj, k = (u64)ck;
                const u64 sgn = pos > 0 ? (u64)1 : ~(u64)0;        // +1 / -1 (mod 2^64)
                d[0] += sgn;
                d[1] += sgn * i; d[2] += sgn * j; d[3] += sgn * k;
                d[4] += sgn * i * i; d[5] += sgn * j * j; d[6] += sgn * k * k;
            }
            if (zer) d[7] += zer > 0 ? (u64)1 : ~(u64)0;
            u[l] = nv;
        }
        flush_stats(d, item, stats);
    }
}

int band_update(const Geom& g, double* u, const double* vel, const i64* band_idx,
                const i64* n_band, double step, double* new_u, double* gmag_out, int apply,
                u64* stats, cudaStream_t s) {
    BandGeom bg;
    bg.ndim = g.ndim; bg.n0 = g.n[0]; bg.n1 = g.n[1]; bg.n2 = g.n[2]; bg.voxels = g.voxels;
    bg.dx0 = g.dx[0]; bg.dx1 = g.dx[1]; bg.dx2 = g.dx[2];
    bg.r0 = g.rdx[0]; bg.r1 = g.rdx[1]; bg.r2 = g.rdx[2];
    bg.inv_voxels = 1.0 / (double)g.voxels;
    bg.small = g.voxels < ((i64)1 << 32) ? 1 : 0;
    bg.single = g.batch == 1 ? 1 : 0;
    band_update_kernel<<<LSML_SMS * 8, 256, 0, s>>>(u, vel, band_idx, n_band, bg, step, new_u, gmag_out);
    LSML_LAUNCH_CHECK("band_update_kernel");
    if (apply) {
        band_apply_kernel<<<LSML_SMS * 4, APPLY_THREADS, 0, s>>>(u, new_u, band_idx, n_band, bg, stats);
        LSML_LAUNCH_CHECK("band_apply_kernel");
    }
    return 0;
}

}  // namespace lsml
