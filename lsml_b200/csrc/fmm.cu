// fmm.cu -- narrow-band signed distance / band rebuild on the device
// (lsml/util/distance_transform.py:5-59 -> skfmm.distance).
//
// The reference rebuilds the band every iteration with scikit-fmm's fast-marching method: a
// sequential heap algorithm.  What has to be reproduced is its RESULT: the frozen set
// {|d| <= band} (bit-exact) and the distances.  The fast-marching solution has a local
// characterisation that a parallel machine can iterate on:
//
//   the value with which voxel x is frozen is the last value the marcher gave it, i.e. the
//   upwind update U(x; S) evaluated on the set S of stencil voxels frozen BEFORE x; voxels
//   are frozen in increasing (|d|, flat index) order.
//
// Given current estimates T(y) of the 4*ndim stencil voxels of x (two each way per axis), V(x)
// replays the marcher from x's point of view: stencil voxels are "frozen" one at a time in key
// order, after each one the update U is re-evaluated exactly as the marcher would (direct
// neighbour frozen -> update; voxel two cells away frozen while the one in between is frozen ->
// update), and the replay stops as soon as x's own key is smaller than the next stencil key
// (x is frozen first).  The marcher's final state is a fixed point of T <- V(T); iterating V
// over the candidate voxels (chaotic relaxation, in place) reaches it in about as many sweeps
// as the band is voxels wide, because the voxels that are final stay final (induction over the
// marcher's freeze order).  Every U uses the marcher's own floating-point expressions in the
// same order (no FMA), so the converged values are bit-identical to the sequential result,
// not merely close.  oracle/lsml_oracle.c:orc_fmm_distance is the sequential statement.
//
// Kernels:
//   fmm_clear_kernel  reset T / mask of the voxels touched by the previous rebuild (sparse)
//   fmm_init_kernel   dense pass over u: exact zeros and sign-change voxels are frozen with the
//                     sub-voxel interpolated distance (8 B/voxel read: HBM-bound)
//   fmm_seed_kernel   direct neighbours of the initial frozen set become candidates
//   fmm_march_kernel  cooperative persistent kernel: sweeps of V over the candidates with a
//                     grid barrier between sweeps, candidate set grown from voxels whose value
//                     is inside the band; then the final mask
//
// T is +inf for every voxel that has no value (the dense array is filled once and only the
// touched voxels are reset), mask doubles as the state array: 0 far, 1 initial frozen,
// 2 candidate during the march; 0/1 = the reference's `mask` afterwards.
#include <cooperative_groups.h>
#include <cfloat>
#include "lsml_internal.h"

namespace cg = cooperative_groups;

namespace lsml {

struct FmmGeom {
    int n0, n1, n2;
    i64 voxels, batch;
    double dx[3];     // padded
    double idx2[3];   // 1/dx/dx
};

struct FmmCounters {      // device-resident, 8 x i64
    i64 n_init;           // initial frozen voxels: list[0 .. n_init)
    i64 n_cand;           // candidates: list[cap-1 .. cap-n_cand] (growing downwards)
    i64 changed;          // per-sweep change flag
    i64 sweeps;           // sweeps executed by the last march
    i64 converged;        // 1 when the last march reached its fixed point
    i64 prev_init, prev_cand;   // extents to clear at the next rebuild
    i64 pad;
    i64 flags[4];         // rotating per-sweep "changed" flags of the march kernel
};

// Ordering decisions are made on |d| rounded to float32 (see oracle/lsml_oracle.c:qmag): values
// that tie in exact arithmetic but differ by float64 rounding noise tie exactly here and are
// ordered by flat index.
__device__ __forceinline__ float qmag(double d) { return __double2float_rn(fabs(d)); }

__device__ __forceinline__ double pos_inf() { return __longlong_as_double(0x7ff0000000000000LL); }

// State byte of a voxel during the march: 0 far, 1 initial frozen, 2 candidate (clean),
// 6 candidate whose stencil changed since it was last evaluated (dirty).
constexpr u8 ST_INIT = 1, ST_CAND = 2, ST_DIRTY = 6;

// claim voxel g (state 0 -> 6); returns true for the single winner.  The state is one byte of
// a 32-bit word; atomicOr on the word keeps the other three bytes intact.
__device__ __forceinline__ bool claim(u8* mask, i64 g) {
    unsigned* word = reinterpret_cast<unsigned*>(mask + (g & ~(i64)3));
    const unsigned shift = 8u * (unsigned)(g & 3);
    const unsigned old = atomicOr(word, (unsigned)ST_DIRTY << shift);
    return ((old >> shift) & 0xffu) == 0u;
}

__global__ void __launch_bounds__(256)
fmm_clear_kernel(double* __restrict__ T, u8* __restrict__ mask, const i64* __restrict__ list,
                 i64 cap, FmmCounters* ctr) {
    const i64 ni = ctr->prev_init, nc = ctr->prev_cand;
    for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < ni + nc;
         t += (i64)gridDim.x * blockDim.x) {
        const i64 g = t < ni ? list[t] : list[cap - 1 - (t - ni)];
        T[g] = pos_inf();
        mask[g] = 0;
    }
}

__global__ void fmm_reset_counters_kernel(FmmCounters* ctr) {
    ctr->n_init = 0; ctr->n_cand = 0; ctr->changed = 0; ctr->sweeps = 0; ctr->converged = 0;
}

// Initial frozen set (orc_fmm_distance "initial frozen set"): zeros, and voxels with an axis
// neighbour of opposite sign; d = sign / sqrt(sum_a 1/d_a^2), d_a = nearer linear crossing.
// Items that are single-signed (no positive or no non-positive... see distance_transform.py:42)
// are left untouched => empty mask.  stats[item*8+0] = #(u>0), stats[item*8+7] = #(u==0).
template <int NDIM>
__global__ void __launch_bounds__(256)
fmm_init_kernel(const double* __restrict__ u, double* __restrict__ T, u8* __restrict__ mask,
                i64* __restrict__ list, FmmGeom g, const u64* __restrict__ stats,
                FmmCounters* ctr) {
    // blockIdx.x = item*n0 + i, blockIdx.y = j tile, blockIdx.z = k tile (no per-thread division)
    const int k = blockIdx.z * blockDim.x + threadIdx.x;
    const int cj = blockIdx.y * blockDim.y + threadIdx.y;
    if (k >= g.n2 || cj >= g.n1) return;
    const i64 item = blockIdx.x / (unsigned)g.n0;
    const int ci = (int)(blockIdx.x % (unsigned)g.n0);
    const u64 npos = stats[item * 8 + 0], nzero = stats[item * 8 + 7];
    if (nzero != (u64)g.voxels && (npos == 0 || npos == (u64)g.voxels)) return;
    const i64 l = ((i64)blockIdx.x * g.n1 + cj) * g.n2 + k;
    const double c = __ldg(u + l);
    double d;
    if (c == 0.0) {
        d = 0.0;
    } else {
        const int coord[3] = {ci, cj, k};
        const int ext[3] = {g.n0, g.n1, g.n2};
        const i64 st[3] = {(i64)g.n1 * g.n2, (i64)g.n2, 1};
        bool borders = false;
        double dsum = 0.0;
#pragma unroll
        for (int ax = 3 - NDIM; ax < 3; ++ax) {
            double ld = 0.0;
#pragma unroll
            for (int j = -1; j < 2; j += 2) {
                const int cc = coord[ax] + j;
                if (cc < 0 || cc >= ext[ax]) continue;
                const double nv = __ldg(u + l + j * st[ax]);
                if (c * nv < 0) {
                    borders = true;
                    const double dd = g.dx[ax] * c / (c - nv);
                    if (ld == 0 || ld > dd) ld = dd;
                }
            }
            if (ld > 0) dsum += 1 / ld / ld;
        }
        if (!borders) return;
        d = c < 0 ? -sqrt(1 / dsum) : sqrt(1 / dsum);
    }
    T[l] = d;
    mask[l] = 1;
    const i64 pos = (i64)atomicAdd(reinterpret_cast<u64*>(&ctr->n_init), (u64)1);
    list[pos] = l;
}

// voxel coordinates (padded axes) of flat index l
__device__ __forceinline__ void coords_of(i64 l, const FmmGeom& g, int coord[3]) {
    const i64 loc = l % g.voxels;
    coord[2] = (int)(loc % g.n2);
    const i64 r = loc / g.n2;
    coord[1] = (int)(r % g.n1);
    coord[0] = (int)(r / g.n1);
}

template <int NDIM>
__device__ __forceinline__ void push_neighbours(i64 l, const FmmGeom& g, double* T, u8* mask,
                                                i64* list, i64 cap, FmmCounters* ctr) {
    int coord[3];
    coords_of(l, g, coord);
    const int ext[3] = {g.n0, g.n1, g.n2};
    const i64 st[3] = {(i64)g.n1 * g.n2, (i64)g.n2, 1};
#pragma unroll
    for (int ax = 3 - NDIM; ax < 3; ++ax) {
#pragma unroll
        for (int j = -1; j < 2; j += 2) {
            const int cc = coord[ax] + j;
            if (cc < 0 || cc >= ext[ax]) continue;
            const i64 n = l + j * st[ax];
            if (__ldcg(mask + n) == 0 && claim(mask, n)) {
                const i64 pos = (i64)atomicAdd(reinterpret_cast<u64*>(&ctr->n_cand), (u64)1);
                list[cap - 1 - pos] = n;
            }
        }
    }
}

template <int NDIM>
__global__ void __launch_bounds__(256)
fmm_seed_kernel(double* T, u8* mask, i64* list, i64 cap, FmmGeom g, FmmCounters* ctr) {
    const i64 ni = ctr->n_init;
    for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < ni;
         t += (i64)gridDim.x * blockDim.x)
        push_neighbours<NDIM>(list[t], g, T, mask, list, cap, ctr);
}

// ---- the marcher's upwind update on a local stencil ------------------------------------------
// t1/t2: values one / two cells away, index [ax][dir] with dir 0 = -1, 1 = +1;
// f1/f2: "frozen now" predicates.  Same expressions, in the same order, as the oracle's
// fmm_update()/fmm_solve() (oracle/lsml_oracle.c), including the causality rule: a root that is
// smaller in magnitude than a value it was computed from (or a non-positive discriminant)
// removes the contributor with the largest magnitude and solves again.
template <int NDIM>
__device__ __forceinline__ double stencil_update(const double (&t1)[3][2], const double (&t2)[3][2],
                                                 const bool (&f1)[3][2], const bool (&f2)[3][2],
                                                 const FmmGeom& g, double phi) {
    const double aa = 9.0 / 4.0, one_third = 1.0 / 3.0;
    double v1[3], v2[3];
    bool active[3];
    int nactive = 0;
#pragma unroll
    for (int ax = 3 - NDIM; ax < 3; ++ax) {
        v1[ax] = DBL_MAX; v2[ax] = DBL_MAX;
#pragma unroll
        for (int d = 0; d < 2; ++d) {
            if (f1[ax][d] && (v1[ax] == DBL_MAX || qmag(t1[ax][d]) < qmag(v1[ax]))) {
                v1[ax] = t1[ax][d];
                v2[ax] = DBL_MAX;
                if (f2[ax][d] && (v1[ax] == 0.0 || t2[ax][d] * v1[ax] < 0 ||
                                  qmag(t2[ax][d]) <= qmag(v1[ax])))
                    v2[ax] = t2[ax][d];
            }
        }
        active[ax] = v1[ax] < DBL_MAX;
        nactive += active[ax] ? 1 : 0;
    }
#pragma unroll 1
    while (nactive > 0) {
        double a = 0, b = 0, c = 0;
#pragma unroll
        for (int ax = 3 - NDIM; ax < 3; ++ax) {
            if (!active[ax]) continue;
            if (v2[ax] < DBL_MAX) {
                const double tp = one_third * (4 * v1[ax] - v2[ax]);
                a += g.idx2[ax] * aa;
                b -= g.idx2[ax] * 2 * aa * tp;
                c += g.idx2[ax] * aa * (tp * tp);
            } else {
                a += g.idx2[ax];
                b -= g.idx2[ax] * 2 * v1[ax];
                c += g.idx2[ax] * (v1[ax] * v1[ax]);
            }
        }
        c -= 1;
        const double det = b * b - 4 * a * c;
        if (det > 0) {
            const double r = (phi > DBL_EPSILON) ? (-b + sqrt(det)) / 2.0 / a
                                                 : (-b - sqrt(det)) / 2.0 / a;
            bool causal = true;
#pragma unroll
            for (int ax = 3 - NDIM; ax < 3; ++ax) {
                if (!active[ax]) continue;
                if (qmag(r) <= qmag(v1[ax])) causal = false;
                if (v2[ax] < DBL_MAX && qmag(r) <= qmag(v2[ax])) causal = false;
            }
            if (causal) return r;
        }
        // remove the contributor farthest from the interface (first maximum in the order
        // ax0.v2, ax0.v1, ax1.v2, ...): a value2 demotes its axis, a value1 drops it
        int worst = -1;
        bool worst_is_v2 = false;
        float wv = -1.0f;
#pragma unroll
        for (int ax = 3 - NDIM; ax < 3; ++ax) {
            if (!active[ax]) continue;
            if (v2[ax] < DBL_MAX && qmag(v2[ax]) > wv) { wv = qmag(v2[ax]); worst = ax; worst_is_v2 = true; }
            if (qmag(v1[ax]) > wv) { wv = qmag(v1[ax]); worst = ax; worst_is_v2 = false; }
        }
#pragma unroll
        for (int ax = 3 - NDIM; ax < 3; ++ax)
            if (ax == worst) {
                if (worst_is_v2) v2[ax] = DBL_MAX;
                else active[ax] = false;
            }
        if (!worst_is_v2) nactive--;
    }
    return 0.0;
}

__device__ __forceinline__ bool key_less(double ta, i64 ia, double tb, i64 ib) {
    const float a = qmag(ta), b = qmag(tb);
    return (a < b) || (a == b && ia < ib);
}

// V(x): replay of the marcher around x (see the header).  Returns +inf when x never gets a value.
template <int NDIM>
__device__ double replay(i64 l, const double* __restrict__ u, const double* T, const u8* mask,
                         const FmmGeom& g, double band) {
    int coord[3];
    coords_of(l, g, coord);
    const int ext[3] = {g.n0, g.n1, g.n2};
    const i64 st[3] = {(i64)g.n1 * g.n2, (i64)g.n2, 1};
    double t1[3][2], t2[3][2];
    bool f1[3][2], f2[3][2];          // frozen now
    bool e1[3][2], e2[3][2];          // pending event (marched voxel with a value inside the band)
    const double inf = pos_inf();
#pragma unroll
    for (int ax = 0; ax < 3; ++ax)
#pragma unroll
        for (int d = 0; d < 2; ++d) {
            t1[ax][d] = inf; t2[ax][d] = inf;
            f1[ax][d] = f2[ax][d] = e1[ax][d] = e2[ax][d] = false;
        }
    bool any_direct_frozen = false;
#pragma unroll
    for (int ax = 3 - NDIM; ax < 3; ++ax) {
#pragma unroll
        for (int d = 0; d < 2; ++d) {
            const int j = d ? 1 : -1;
            const int c1 = coord[ax] + j, c2 = coord[ax] + 2 * j;
            if (c1 >= 0 && c1 < ext[ax]) {
                const i64 n = l + j * st[ax];
                const double t = __ldcg(T + n);       // L2: other CTAs update T concurrently
                const u8 s = __ldcg(mask + n);
                t1[ax][d] = t;
                f1[ax][d] = (s == ST_INIT);
                e1[ax][d] = (s & ST_CAND) && (fabs(t) <= band);
                any_direct_frozen |= f1[ax][d];
            }
            if (c2 >= 0 && c2 < ext[ax]) {
                const i64 n = l + 2 * j * st[ax];
                const double t = __ldcg(T + n);
                const u8 s = __ldcg(mask + n);
                t2[ax][d] = t;
                f2[ax][d] = (s == ST_INIT);
                e2[ax][d] = (s & ST_CAND) && (fabs(t) <= band);
            }
        }
    }
    const double phi = __ldg(u + l);
    bool narrow = false;
    double t = inf;
    if (any_direct_frozen) {
        const double d0 = stencil_update<NDIM>(t1, t2, f1, f2, g, phi);
        if (d0 != 0.0) { t = d0; narrow = true; }
    }
    // at most 4*NDIM events
#pragma unroll 1
    for (int it = 0; it < 4 * NDIM; ++it) {
        // next event: smallest key among pending stencil voxels
        int bax = -1, bd = 0, bwhich = 0;
        double bt = inf; i64 bi = 0;
#pragma unroll
        for (int ax = 3 - NDIM; ax < 3; ++ax)
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                const int j = d ? 1 : -1;
                if (e1[ax][d]) {
                    const i64 n = l + j * st[ax];
                    if (bax < 0 || key_less(t1[ax][d], n, bt, bi)) { bax = ax; bd = d; bwhich = 1; bt = t1[ax][d]; bi = n; }
                }
                if (e2[ax][d]) {
                    const i64 n = l + 2 * j * st[ax];
                    if (bax < 0 || key_less(t2[ax][d], n, bt, bi)) { bax = ax; bd = d; bwhich = 2; bt = t2[ax][d]; bi = n; }
                }
            }
        if (bax < 0) break;
        if (narrow && key_less(t, l, bt, bi)) break;      // x is frozen before that voxel
        bool reeval;
#pragma unroll
        for (int ax = 3 - NDIM; ax < 3; ++ax)
#pragma unroll
            for (int d = 0; d < 2; ++d)
                if (ax == bax && d == bd) {
                    if (bwhich == 1) { e1[ax][d] = false; f1[ax][d] = true; }
                    else { e2[ax][d] = false; f2[ax][d] = true; }
                }
        if (bwhich == 1) {
            reeval = true;                                 // direct neighbour frozen
        } else {
            bool mid = false;                              // two cells away: needs the voxel in
#pragma unroll                                             // between frozen and x already narrow
            for (int ax = 3 - NDIM; ax < 3; ++ax)
#pragma unroll
                for (int d = 0; d < 2; ++d)
                    if (ax == bax && d == bd) mid = f1[ax][d];
            reeval = mid && narrow;
        }
        if (reeval) {
            const double dn = stencil_update<NDIM>(t1, t2, f1, f2, g, phi);
            if (dn != 0.0) { t = dn; narrow = true; }
        }
    }
    return narrow ? t : inf;
}

constexpr int FMM_MAX_SWEEPS = 4096;

// every candidate that has x in its stencil (one or two cells along an axis) must look again
template <int NDIM>
__device__ __forceinline__ void mark_stencil_dirty(i64 l, const FmmGeom& g, u8* mask) {
    int coord[3];
    coords_of(l, g, coord);
    const int ext[3] = {g.n0, g.n1, g.n2};
    const i64 st[3] = {(i64)g.n1 * g.n2, (i64)g.n2, 1};
#pragma unroll
    for (int ax = 3 - NDIM; ax < 3; ++ax) {
#pragma unroll
        for (int j = -2; j <= 2; ++j) {
            if (j == 0) continue;
            const int cc = coord[ax] + j;
            if (cc < 0 || cc >= ext[ax]) continue;
            const i64 n = l + j * st[ax];
            if (__ldcg(mask + n) & ST_CAND) __stcg(mask + n, ST_DIRTY);   // unconditional re-set
        }
    }
}

// Sweeps of the replay operator over the candidate list until nothing changes.  Only DIRTY
// candidates are evaluated: a candidate becomes dirty when it is created and whenever a voxel of
// its stencil changes value, so after the first sweep the work follows the front outwards
// instead of touching the whole band every time.  One grid barrier per sweep; the "something
// changed" flags rotate through three slots so that no second barrier is needed to reset them.
// Ordering of the dirty protocol (per voxel): clear own flag -> fence -> read stencil; a writer
// does: store T -> fence -> set the readers' flags.  Either the reader sees the new T, or its
// flag is set again after it cleared it.
template <int NDIM>
__global__ void __launch_bounds__(256)
fmm_march_kernel(const double* __restrict__ u, double* T, u8* mask, i64* list, i64 cap,
                 FmmGeom g, double band, FmmCounters* ctr, i64* flags /* [3] */) {
    cg::grid_group grid = cg::this_grid();
    const i64 tid = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const i64 nthreads = (i64)gridDim.x * blockDim.x;
    volatile i64* vflags = flags;
    volatile i64* vcand = &ctr->n_cand;
    int sweep = 0;
    bool converged = false;
    i64 nc = *vcand;
    for (; sweep < FMM_MAX_SWEEPS; ++sweep) {
        const int slot = sweep % 3;
        if (tid == 0) vflags[(sweep + 1) % 3] = 0;
        bool changed = false;
        for (i64 t = tid; t < nc; t += nthreads) {
            const i64 l = list[cap - 1 - t];
            if (__ldcg(mask + l) != ST_DIRTY) continue;
            __stcg(mask + l, ST_CAND);
            __threadfence();
            const double old = __ldcg(T + l);
            const double nw = replay<NDIM>(l, u, T, mask, g, band);
            if (__double_as_longlong(old) != __double_as_longlong(nw)) {
                __stcg(T + l, nw);
                __threadfence();
                changed = true;
                mark_stencil_dirty<NDIM>(l, g, mask);
                // a voxel whose value enters the band exposes its direct neighbours
                if (fabs(nw) <= band && !(fabs(old) <= band))
                    push_neighbours<NDIM>(l, g, T, mask, list, cap, ctr);
            }
        }
        if (changed) vflags[slot] = 1;
        __threadfence();
        grid.sync();
        const i64 nc_new = *vcand;
        const bool any = (vflags[slot] != 0) || (nc_new != nc);
        nc = nc_new;
        if (!any) { converged = true; ++sweep; break; }
    }
    // final mask: candidates frozen by the marcher are those with |d| <= band
    for (i64 t = tid; t < nc; t += nthreads) {
        const i64 l = list[cap - 1 - t];
        if (fabs(__ldcg(T + l)) <= band) mask[l] = 1;
        else { mask[l] = 0; T[l] = pos_inf(); }
    }
    if (tid == 0) {
        ctr->sweeps = sweep;
        ctr->converged = converged ? 1 : 0;
        ctr->prev_init = ctr->n_init;
        ctr->prev_cand = nc;
        vflags[0] = 0; vflags[1] = 0; vflags[2] = 0;
    }
}

// dist_out = T where mask else 0  (skfmm masked data after post-processing)
__global__ void __launch_bounds__(256)
fmm_export_kernel(const double* __restrict__ T, const u8* __restrict__ mask,
                  double* __restrict__ dist, i64 total) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (i64)gridDim.x * blockDim.x)
        dist[i] = mask[i] ? T[i] : 0.0;
}

__global__ void __launch_bounds__(256)
fill_inf_kernel(double* __restrict__ T, i64 total) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (i64)gridDim.x * blockDim.x)
        T[i] = pos_inf();
}

// ---------------------------------------------------------------------------------- host side
static int g_march_blocks[4] = {0, 0, 0, 0};

template <int NDIM>
static int march_launch(const double* u, double* T, u8* mask, i64* list, i64 cap, FmmGeom fg,
                        double band, FmmCounters* ctr, cudaStream_t s) {
    if (g_march_blocks[NDIM] == 0) {
        int per_sm = 0, dev = 0, sms = 0;
        LSML_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fmm_march_kernel<NDIM>, 256, 0));
        LSML_CUDA(cudaGetDevice(&dev));
        LSML_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (per_sm < 1) { set_error("fmm_march_kernel does not fit on an SM"); return 1; }
        if (per_sm > 2) per_sm = 2;
        g_march_blocks[NDIM] = per_sm * sms;
    }
    i64* flags = ctr->flags;
    void* args[] = {(void*)&u, (void*)&T, (void*)&mask, (void*)&list, (void*)&cap, (void*)&fg,
                    (void*)&band, (void*)&ctr, (void*)&flags};
    LSML_CUDA(cudaLaunchCooperativeKernel((void*)fmm_march_kernel<NDIM>, dim3(g_march_blocks[NDIM]),
                                          dim3(256), args, 0, s));
    count_launch();
    return 0;
}

i64 fmm_workspace_bytes(i64 total) {
    // T (f64) + list (i64) + counters
    return total * (i64)(sizeof(double) + sizeof(i64)) + (i64)sizeof(FmmCounters);
}

// first use of a workspace: T = +inf everywhere, counters zero
int fmm_workspace_init(i64 total, void* workspace, cudaStream_t s) {
    double* T = static_cast<double*>(workspace);
    FmmCounters* ctr = reinterpret_cast<FmmCounters*>(static_cast<char*>(workspace) +
                                                      total * (sizeof(double) + sizeof(i64)));
    fill_inf_kernel<<<LSML_SMS * 8, 256, 0, s>>>(T, total);
    LSML_LAUNCH_CHECK("fill_inf_kernel");
    LSML_CUDA(cudaMemsetAsync(ctr, 0, sizeof(FmmCounters), s));
    return 0;
}

// Rebuild the band of every item from u.  mask must be all-zero on first use of the workspace
// and is thereafter maintained by this function.  band == 0 => whole grid (no narrow band).
// stats: the exact {u>0} statistics of the SAME u (global_stats), used for the reference's
// single-sign / all-zero edge cases.  dist_out may be null.
int fmm_rebuild(const Geom& g, const double* u, double band, const u64* stats, u8* mask,
                double* dist_out, void* workspace, cudaStream_t s) {
    const i64 total = g.voxels * g.batch;
    double* T = static_cast<double*>(workspace);
    i64* list = reinterpret_cast<i64*>(static_cast<char*>(workspace) + total * sizeof(double));
    FmmCounters* ctr = reinterpret_cast<FmmCounters*>(static_cast<char*>(workspace) +
                                                      total * (sizeof(double) + sizeof(i64)));
    FmmGeom fg;
    fg.n0 = g.n[0]; fg.n1 = g.n[1]; fg.n2 = g.n[2]; fg.voxels = g.voxels; fg.batch = g.batch;
    for (int a = 0; a < 3; ++a) { fg.dx[a] = g.dx[a]; fg.idx2[a] = 1.0 / g.dx[a] / g.dx[a]; }
    const double eff_band = band != 0.0 ? band : DBL_MAX;

    fmm_clear_kernel<<<LSML_SMS * 8, 256, 0, s>>>(T, mask, list, total, ctr);
    LSML_LAUNCH_CHECK("fmm_clear_kernel");
    fmm_reset_counters_kernel<<<1, 1, 0, s>>>(ctr);
    LSML_LAUNCH_CHECK("fmm_reset_counters_kernel");

    int bx = 32;
    while (bx < 256 && bx < g.n[2]) bx <<= 1;
    int by = 256 / bx;
    if (by > g.n[1]) { by = 1; while (by < g.n[1]) by <<= 1; }
    const dim3 block(bx, by), grid((unsigned)(g.batch * g.n[0]), (unsigned)((g.n[1] + by - 1) / by),
                                   (unsigned)((g.n[2] + bx - 1) / bx));
    int st = 0;
    if (g.ndim == 3) {
        fmm_init_kernel<3><<<grid, block, 0, s>>>(u, T, mask, list, fg, stats, ctr);
        LSML_LAUNCH_CHECK("fmm_init_kernel");
        fmm_seed_kernel<3><<<LSML_SMS * 8, 256, 0, s>>>(T, mask, list, total, fg, ctr);
        LSML_LAUNCH_CHECK("fmm_seed_kernel");
        st = march_launch<3>(u, T, mask, list, total, fg, eff_band, ctr, s);
    } else if (g.ndim == 2) {
        fmm_init_kernel<2><<<grid, block, 0, s>>>(u, T, mask, list, fg, stats, ctr);
        LSML_LAUNCH_CHECK("fmm_init_kernel");
        fmm_seed_kernel<2><<<LSML_SMS * 8, 256, 0, s>>>(T, mask, list, total, fg, ctr);
        LSML_LAUNCH_CHECK("fmm_seed_kernel");
        st = march_launch<2>(u, T, mask, list, total, fg, eff_band, ctr, s);
    } else {
        fmm_init_kernel<1><<<grid, block, 0, s>>>(u, T, mask, list, fg, stats, ctr);
        LSML_LAUNCH_CHECK("fmm_init_kernel");
        fmm_seed_kernel<1><<<LSML_SMS * 8, 256, 0, s>>>(T, mask, list, total, fg, ctr);
        LSML_LAUNCH_CHECK("fmm_seed_kernel");
        st = march_launch<1>(u, T, mask, list, total, fg, eff_band, ctr, s);
    }
    if (st) return st;
    if (dist_out != nullptr) {
        fmm_export_kernel<<<LSML_SMS * 8, 256, 0, s>>>(T, mask, dist_out, total);
        LSML_LAUNCH_CHECK("fmm_export_kernel");
    }
    return 0;
}

// copies {n_init, n_cand, sweeps, converged} of the last rebuild into out[4] (device -> device)
int fmm_counters_ptr(i64 total, void* workspace, const i64** out) {
    *out = reinterpret_cast<const i64*>(static_cast<char*>(workspace) +
                                        total * (sizeof(double) + sizeof(i64)));
    return 0;
}

}  // namespace lsml
