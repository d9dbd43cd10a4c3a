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
// (x is frozen first).  The marcher's final state is the UNIQUE fixed point of T <- V(T)
// (induction over the marcher's freeze order); iterating V over the candidate voxels in any
// order (chaotic relaxation, in place) reaches it in about as many sweeps as the marcher's
// dependency graph is deep.  Every U uses the marcher's own floating-point expressions in the
// same order (no FMA), so the converged values are bit-identical to the sequential result,
// not merely close.  oracle/lsml_oracle.c:orc_fmm_distance is the sequential statement,
// oracle/fmm_replay_model.c the CPU model of V (tests/test_fmm_algorithm.py).
//
// Kernels of one rebuild:
//   fmm_clear_kernel      reset T / state / epoch of the voxels the previous rebuild touched
//   fmm_prepare_kernel    counters, hint generation, parking-bucket offsets
//   fmm_init_kernel       dense search of the interface: exact zeros and sign-change voxels are
//                         frozen with the sub-voxel interpolated distance (first rebuild of a
//                         level set; boundary / ghost planes of a slab)
//   fmm_init_band_kernel  the same search restricted to the previous band (evolution loop)
//   fmm_seed_kernel       direct neighbours of the initial frozen set become candidates
//   fmm_march_kernel      cooperative persistent kernel: sweeps of V over the queued candidates
//                         (work ring + parking buckets), one grid barrier per sweep, small work
//                         sets on a single CTA; then the final mask
//   fmm_requeue_kernel, fmm_finalize_kernel, fmm_pending_kernel   slab decomposition only
//
// Per-voxel state: T (f64, +inf = no value; the dense array is filled once and only touched
// voxels are ever reset), Aux (8 B: queue epoch, reach, depth hint), and the mask byte, which
// doubles as the state during a rebuild: 0 far, 1 initial frozen, (depth << 2) | 2 candidate;
// 0/1 = the reference's `mask` afterwards.
#include <cooperative_groups.h>
#include <cooperative_groups/scan.h>
#include <cfloat>
#include <cstddef>
#include <cstdlib>
#include "lsml_internal.h"

namespace cg = cooperative_groups;

namespace lsml {

// [host-test:begin] (tests/test_fmm_device_code_host.py compiles the marked regions with g++)
struct FmmGeom {
    int n0, n1, n2;
    i64 voxels, batch;
    double dx[3];     // padded
    double idx2[3];   // 1/dx/dx
    // list entries are "packed coordinates" (item, i, j, k) in bit fields: k at bit 0, j at sh1,
    // i at sh0, item at shb -- decoding is shifts and masks, never a 64-bit division
    int sh1, sh0, shb;
    // slab decomposition: this rank OWNS the planes [own_lo, own_hi) of (padded) axis 0 and only
    // ever freezes / queues / evaluates voxels there; the other planes of the local grid are
    // halo copies of the neighbour ranks' voxels (whole grid = [0, n0) for a single GPU)
    int own_lo, own_hi;
};
// [host-test:end]

static int bits_for(i64 n) { int b = 0; while (((i64)1 << b) < n) ++b; return b; }

static int make_fmm_geom(const Geom& g, FmmGeom& fg) {
    fg.n0 = g.n[0]; fg.n1 = g.n[1]; fg.n2 = g.n[2]; fg.voxels = g.voxels; fg.batch = g.batch;
    for (int a = 0; a < 3; ++a) { fg.dx[a] = g.dx[a]; fg.idx2[a] = 1.0 / g.dx[a] / g.dx[a]; }
    fg.sh1 = bits_for(g.n[2]);
    fg.sh0 = fg.sh1 + bits_for(g.n[1]);
    fg.shb = fg.sh0 + bits_for(g.n[0]);
    if (fg.shb + bits_for(g.batch) > 62) { set_error("grid too large for packed band entries"); return 1; }
    fg.own_lo = 0; fg.own_hi = g.n[0];
    return 0;
}

// [host-sched:begin] (tests/test_fmm_scheduler_host.py: the scheduler's device functions on one host thread)
struct FmmCounters {      // device-resident
    i64 n_init;           // initial frozen voxels: list[0 .. n_init)
    i64 n_cand;           // candidates: list[cap-1 .. cap-n_cand] (growing downwards)
    i64 tail_sweeps;      // sweeps of the last march that ran on one CTA (diagnostic)
    i64 sweeps;           // sweeps executed by the last march
    i64 converged;        // 1 when the last march reached its fixed point
    i64 prev_init, prev_cand;   // extents to clear at the next rebuild
    i64 replays;          // replay evaluations of the last march (diagnostic)
    // ---- work ring: packed entries of the candidates queued for the next sweep, in push order
    i64 tail3[3];         // entries queued for the r-th executed sweep are counted in tail3[r % 3]:
                          // stable while that sweep reads it, zeroed during the next one, refilled
                          // during the one after (no reader can race a writer, no extra barrier)
    i64 pub_begin, pub_end;     // window published by CTA 0 after a run of single-CTA sweeps
    i64 pub_sweep, pub_rix;     // ... and the sweep number / executed-sweep count reached
    i64 resume_sweep, resume_rix, resume_begin;   // where the next march launch of this rebuild
                                // continues (slab decomposition: several launches per rebuild)
    i64 overflow[3];      // rotating "a push was dropped" flags => next sweep scans the list
    // ---- schedule hints carried from one rebuild to the next (see Aux::lvl)
    i64 gen;              // rebuild counter (low 8 bits tag the hints)
    i64 use_hints;        // this rebuild parks candidates until the sweep of their hint
    i64 last_bucket;      // largest sweep number anything is parked for
    i64 hist[2][256];     // [gen & 1][level]: voxels whose hint of that rebuild is `level`
    i64 offs[257];        // start of each level's bucket in the parking array (scan of hist[prev])
    i64 fill[256];        // entries parked per level so far
    i64 sizes[32];        // work-set size of the first 32 grid-wide sweeps (diagnostic)
    i64 stamps[32];       // %globaltimer (ns) at the start of those sweeps (diagnostic)
    // ---- status the host layer reads ONCE per segmentation (never reset by a rebuild)
    i64 failed_rebuilds;  // marches that ended without reaching their fixed point (sweep limits)
    i64 barriers;         // grid barriers of the last march (diagnostic)
    i64 live[4];          // {sweep, rix, begin, end} published by thread 0 before every grid-wide
                          // step: readable from another stream while a march is running (hang
                          // diagnosis, tools/c4_probe.py)
    i64 hint_slack;       // s > 0: hints are written as depth + (depth >> s) (batches, see
                          // process_candidate); 0: the depth itself
    i64 phase[8];         // -DFMM_PHASE_CLOCKS: SM cycles one lane spent per phase of the small
                          // grid-wide sweeps {loop top -> entry, replay, record, publish, fence,
                          // barrier, -, number of sweeps sampled} (diagnostic, tools/fmm_probe.py)
};
// [host-sched:end]

// [host-test:begin] (tests/test_fmm_device_code_host.py compiles the marked regions with g++)
// Ordering decisions are made on |d| rounded to float32 (see oracle/lsml_oracle.c:qmag): values
// that tie in exact arithmetic but differ by float64 rounding noise tie exactly here and are
// ordered by flat index.
__device__ __forceinline__ float qmag(double d) { return __double2float_rn(fabs(d)); }

__device__ __forceinline__ double pos_inf() { return __longlong_as_double(0x7ff0000000000000LL); }

// State byte of a voxel during the march: 0 far, 1 initial frozen, (level << 2) | 2 candidate,
// where level is the candidate's dependency depth in this rebuild (see Aux::lvl).
constexpr u8 ST_INIT = 1, ST_CAND = 2;
constexpr unsigned MAX_LEVEL = 63u;
// [host-test:end]

// [host-sched:begin] (tests/test_fmm_scheduler_host.py: the scheduler's device functions on one host thread)
// claim voxel g (state 0 -> `value`); returns true for the single winner.  The state is one
// byte of a 32-bit word; atomicOr on the word keeps the other three bytes intact.
__device__ __forceinline__ bool claim(u8* mask, i64 g, unsigned value) {
    unsigned* word = reinterpret_cast<unsigned*>(mask + (g & ~(i64)3));
    const unsigned shift = 8u * (unsigned)(g & 3);
    const unsigned old = atomicOr(word, value << shift);
    return ((old >> shift) & 0xffu) == 0u;
}

// Scheduling record of a voxel (8 bytes, next to T):
//   epoch  the last sweep this voxel was queued for (0 = never); voxel y is in the work set of
//          sweep s exactly when epoch == s, so queueing is an atomicMax and needs no clearing;
//   reach  how far y's last replay looked: max(magnitude of the last event it consumed,
//          magnitude of its result), +inf when it ended without a value, rounded UP to the top
//          16 bits of a float32.  A stencil voxel whose old AND new magnitudes are both beyond
//          `reach` was not consumed by that replay and would not be consumed by a new one
//          (every event before it is unchanged and the replay stops at it), so its change
//          cannot alter y -- y is not queued for it.
//   lvl    schedule hint: (rebuild tag << 8) | dependency depth of y in that rebuild, i.e.
//          1 + the largest depth among the stencil voxels its last replay consumed (initial
//          frozen voxels have depth 0).  A voxel of depth d cannot have its final value before
//          sweep d, and has it at sweep d if everything shallower was evaluated on time.  The
//          dependency structure barely moves between two evolution iterations, so in the NEXT
//          rebuild y is parked until sweep d instead of being re-evaluated every time one of
//          its still-unsettled upwind voxels moves.  A hint only changes WHEN y is evaluated,
//          never the fixed point the sweeps converge to.
struct __align__(8) Aux {
    unsigned epoch;
    unsigned short reach16;
    unsigned short lvl;
};

// Hint tag of rebuild `gen`, already shifted into the high byte of Aux::lvl: 1..255, never 0, so
// that lvl == 0 (never evaluated, or wiped) carries no rebuild's tag.  The tags repeat every 255
// rebuilds; all hints are wiped every 128 (fmm_clear_kernel), so a tag is never ambiguous.
__device__ __forceinline__ unsigned hint_tag_of(i64 gen) {
    return (unsigned)((gen % 255) + 1) << 8;
}

__device__ __forceinline__ unsigned short reach_up16(float r) {
    return (unsigned short)((__float_as_uint(r) + 0xffffu) >> 16);
}
__device__ __forceinline__ float reach_from16(unsigned r16) { return __uint_as_float(r16 << 16); }
constexpr unsigned REACH16_INF = 0x7f80u;

__device__ __forceinline__ i64 decode(i64 e, const FmmGeom& g, int coord[3]) {
    coord[2] = (int)(e & (((i64)1 << g.sh1) - 1));
    coord[1] = (int)((e >> g.sh1) & (((i64)1 << (g.sh0 - g.sh1)) - 1));
    coord[0] = (int)((e >> g.sh0) & (((i64)1 << (g.shb - g.sh0)) - 1));
    const i64 item = e >> g.shb;
    return ((item * g.n0 + coord[0]) * g.n1 + coord[1]) * g.n2 + coord[2];
}
__device__ __forceinline__ i64 encode(i64 item, int i, int j, int k, const FmmGeom& g) {
    return (item << g.shb) | ((i64)i << g.sh0) | ((i64)j << g.sh1) | (i64)k;
}

__global__ void __launch_bounds__(256)
fmm_clear_kernel(double* __restrict__ T, Aux* __restrict__ aux, u8* __restrict__ mask,
                 const i64* __restrict__ list, i64 cap, FmmGeom g, FmmCounters* ctr, i64 total) {
    const i64 ni = ctr->prev_init, nc = ctr->prev_cand;
    const bool wipe = (ctr->gen & 127) == 127;
    for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < ni + nc;
         t += (i64)gridDim.x * blockDim.x) {
        int coord[3];
        const i64 l = decode(t < ni ? list[t] : list[cap - 1 - (t - ni)], g, coord);
        T[l] = pos_inf();
        // the hint survives (except in a wipe: both loops then store 0, in either order)
        aux[l] = Aux{0u, (unsigned short)REACH16_INF, wipe ? (unsigned short)0 : aux[l].lvl};
        mask[l] = 0;
    }
    // Hints are tagged with hint_tag_of(rebuild counter), period 255.  Every 128th rebuild ALL
    // hints are dropped (one dense 8 B/voxel pass), so no voxel can carry a tag that is 255
    // rebuilds old and would be mistaken for last rebuild's: the parking buckets are sized from last
    // rebuild's histogram and must never receive a voxel it did not count.
    if (wipe) {
        for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total;
             i += (i64)gridDim.x * blockDim.x)
            aux[i].lvl = 0;
    }
}

// [host-sched:end]

// Start of a rebuild (one CTA of 256 threads): reset the per-rebuild counters, advance the hint
// generation, turn last rebuild's hint histogram into parking-bucket offsets.
__global__ void __launch_bounds__(256)
fmm_prepare_kernel(FmmCounters* ctr, int allow_hints, i64 park_cap, int hint_slack) {
    __shared__ i64 scan[256];
    const int t = threadIdx.x;
    if (t == 0) {
        ctr->n_init = 0; ctr->n_cand = 0; ctr->tail_sweeps = 0; ctr->sweeps = 0; ctr->converged = 0;
        ctr->replays = 0; ctr->tail3[0] = 0; ctr->tail3[1] = 0; ctr->tail3[2] = 0; ctr->pub_begin = 0; ctr->pub_end = 0; ctr->pub_sweep = 0; ctr->pub_rix = 0;
        ctr->resume_sweep = 1; ctr->resume_rix = 1; ctr->resume_begin = 0;
        ctr->overflow[0] = 0; ctr->overflow[1] = 0; ctr->overflow[2] = 0;
        ctr->last_bucket = 0;
        ctr->hint_slack = hint_slack;
        ctr->gen = ctr->gen + 1;
    }
    __syncthreads();
    const int cur = (int)(ctr->gen & 1), prev = cur ^ 1;
    scan[t] = ctr->hist[prev][t];
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
        const i64 v = t >= o ? scan[t - o] : 0;
        __syncthreads();
        scan[t] += v;
        __syncthreads();
    }
    ctr->offs[t + 1] = scan[t];
    if (t == 0) ctr->offs[0] = 0;
    ctr->hist[cur][t] = 0;
    ctr->fill[t] = 0;
    // every 16th rebuild runs without hints so that they cannot drift upwards for ever
    if (t == 255) ctr->use_hints = (allow_hints && (ctr->gen % 16) != 0 && scan[255] <= park_cap) ? 1 : 0;
}

// [host-test:begin] (tests/test_fmm_device_code_host.py compiles the marked regions with g++)
// Initial frozen value of voxel l (orc_fmm_distance "initial frozen set"): zeros, and voxels
// with an axis neighbour of opposite sign; d = sign / sqrt(sum_a 1/d_a^2), d_a = nearer linear
// crossing.  Returns false when l is not in the initial frozen set.
template <int NDIM>
__device__ __forceinline__ bool init_distance(const double* __restrict__ u, i64 l,
                                              const int (&coord)[3], const FmmGeom& g, double& d) {
    const double c = __ldg(u + l);
    if (c == 0.0) { d = 0.0; return true; }
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
    if (!borders) return false;
    d = c < 0 ? -sqrt(1 / dsum) : sqrt(1 / dsum);
    return true;
}
// [host-test:end]

// [host-sched:begin] (tests/test_fmm_scheduler_host.py: the scheduler's device functions on one host thread)
// Items that are single-signed (no positive or no non-positive... see distance_transform.py:42)
// are left untouched => empty mask.  stats[item*8+0] = #(u>0), stats[item*8+7] = #(u==0).
__device__ __forceinline__ bool item_has_contour(const u64* __restrict__ stats, i64 item,
                                                 const FmmGeom& g) {
    if (stats == nullptr) return true;            // the caller decided (slab decomposition)
    const u64 npos = stats[item * 8 + 0], nzero = stats[item * 8 + 7];
    return nzero == (u64)g.voxels || !(npos == 0 || npos == (u64)g.voxels);
}

// Dense pass over the planes [p_lo, p_hi) of u (first rebuild of a level set, and the two
// boundary planes of a slab on every rebuild): 8 B/voxel read, HBM-bound.
template <int NDIM>
__global__ void __launch_bounds__(256)
fmm_init_kernel(const double* __restrict__ u, double* __restrict__ T, u8* __restrict__ mask,
                i64* __restrict__ list, FmmGeom g, const u64* __restrict__ stats,
                FmmCounters* ctr, int p_lo, int p_hi) {
    // blockIdx.x = item*(p_hi-p_lo) + (i-p_lo), blockIdx.y = j tile, blockIdx.z = k tile
    const int k = blockIdx.z * blockDim.x + threadIdx.x;
    const int cj = blockIdx.y * blockDim.y + threadIdx.y;
    if (k >= g.n2 || cj >= g.n1) return;
    const unsigned np = (unsigned)(p_hi - p_lo);
    const i64 item = blockIdx.x / np;
    const int ci = p_lo + (int)(blockIdx.x % np);
    if (!item_has_contour(stats, item, g)) return;
    const i64 l = (((i64)item * g.n0 + ci) * g.n1 + cj) * g.n2 + k;
    const int coord[3] = {ci, cj, k};
    double d;
    if (!init_distance<NDIM>(u, l, coord, g, d)) return;
    if (!claim(mask, l, ST_INIT)) return;
    T[l] = d;
    const i64 pos = (i64)atomicAdd(reinterpret_cast<u64*>(&ctr->n_init), (u64)1);
    list[pos] = encode(item, ci, cj, k, g);
}

// Sparse variant for the evolution loop: between two rebuilds u changes only on the previous
// band, so a voxel can only be in the new initial frozen set if it is in the previous band or
// is the opposite-signed direct neighbour of a previous-band voxel (a voxel outside the band
// whose neighbours are all outside the band kept its status, and that status was "not frozen").
// One thread per previous-band voxel x evaluates x and every direct neighbour of opposite
// sign; duplicates are resolved by the atomic claim on the state byte.
template <int NDIM>
__global__ void __launch_bounds__(256)
fmm_init_band_kernel(const double* __restrict__ u, double* __restrict__ T, u8* __restrict__ mask,
                     i64* __restrict__ list, FmmGeom g, const u64* __restrict__ stats,
                     const i64* __restrict__ band_idx, const i64* __restrict__ n_band,
                     FmmCounters* ctr) {
    const i64 nb = *n_band;
    const int ext[3] = {g.n0, g.n1, g.n2};
    const i64 st[3] = {(i64)g.n1 * g.n2, (i64)g.n2, 1};
    for (i64 b = (i64)blockIdx.x * blockDim.x + threadIdx.x; b < nb;
         b += (i64)gridDim.x * blockDim.x) {
        const i64 l = __ldg(band_idx + b);
        const i64 item = l / g.voxels, loc = l - item * g.voxels;
        if (!item_has_contour(stats, item, g)) continue;
        int coord[3];
        coord[2] = (int)(loc % g.n2);
        const i64 r = loc / g.n2;
        coord[1] = (int)(r % g.n1);
        coord[0] = (int)(r / g.n1);
        const double c = __ldg(u + l);
        double d;
        if (init_distance<NDIM>(u, l, coord, g, d) && claim(mask, l, ST_INIT)) {
            T[l] = d;
            const i64 pos = (i64)atomicAdd(reinterpret_cast<u64*>(&ctr->n_init), (u64)1);
            list[pos] = encode(item, coord[0], coord[1], coord[2], g);
        }
#pragma unroll
        for (int ax = 3 - NDIM; ax < 3; ++ax) {
#pragma unroll
            for (int j = -1; j < 2; j += 2) {
                const int cc = coord[ax] + j;
                if (cc < 0 || cc >= ext[ax]) continue;
                if (ax == 0 && (cc < g.own_lo || cc >= g.own_hi)) continue;   // the owner does it
                const i64 n = l + j * st[ax];
                if (!(c * __ldg(u + n) < 0)) continue;
                if (__ldcg(mask + n) != 0) continue;             // already claimed
                int nc[3] = {coord[0], coord[1], coord[2]};
                nc[ax] = cc;
                double dn;
                if (init_distance<NDIM>(u, n, nc, g, dn) && claim(mask, n, ST_INIT)) {
                    T[n] = dn;
                    const i64 pos = (i64)atomicAdd(reinterpret_cast<u64*>(&ctr->n_init), (u64)1);
                    list[pos] = encode(item, nc[0], nc[1], nc[2], g);
                }
            }
        }
    }
}

// Work ring: position p lives in slot p % ring_cap.  A sweep consumes the window
// [begin, end) and appends the candidates it dirties behind `end`; an append that would lap the
// window being consumed is dropped and reported (the next sweep then finds the dirty
// candidates by scanning the candidate list instead).
struct Ring {
    i64* slots;
    i64 cap;
};

// Same-address atomics serialise in L2 (about one per clock per address): a sweep that made every
// one of its ~10^5 threads bump the same ring / list / histogram counter spent most of its
// time there.  These helpers let the currently converged lanes of a warp that target the SAME
// counter issue ONE atomic between them.
// returns this lane's slot: (old counter value) + (rank among the lanes sharing `key`)
__device__ __forceinline__ i64 warp_agg_reserve(i64* counter, unsigned key) {
    const unsigned active = __activemask();
    const unsigned peers = __match_any_sync(active, key);
    const unsigned lane = threadIdx.x & 31u;
    const int leader = __ffs(peers) - 1;
    i64 base = 0;
    if ((int)lane == leader)
        base = (i64)atomicAdd(reinterpret_cast<u64*>(counter), (u64)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    return base + __popc(peers & ((1u << lane) - 1u));
}
// counter[key] += delta for every lane, one reduction-atomic per distinct key
__device__ __forceinline__ void warp_agg_add(i64* counters, unsigned key, i64 delta) {
    const unsigned active = __activemask();
    const unsigned peers = __match_any_sync(active, key);
    const unsigned lane = threadIdx.x & 31u;
    if ((int)lane == __ffs(peers) - 1)
        atomicAdd(reinterpret_cast<u64*>(counters + key), (u64)(delta * __popc(peers)));
}

// everything a thread needs to queue work
struct Sched {
    Aux* aux;
    Ring ring;
    i64* park;           // parking array: bucket of level h = park[offs[h] .. offs[h] + fill[h])
    FmmCounters* ctr;
    i64 begin;           // start of the ring window being consumed (overflow test)
    i64 base;            // ring position of the first entry of the sweep being filled
    unsigned next;       // epoch of the sweep being filled
    int tslot;           // its counter: tail3[tslot]
    unsigned hint_tag;   // (tag of the previous rebuild) << 8, or ~0 when hints are off
    int ovf_slot;
    unsigned slack;      // FmmCounters::hint_slack (0 when not given: off)
};

// sweep a voxel with hint word `lvl` should be queued for: the next one, or later if last
// rebuild's hint says its value settled later
__device__ __forceinline__ unsigned target_sweep(const Sched& sc, unsigned lvl) {
    const unsigned h = lvl & 0xffu;
    return ((lvl & 0xff00u) == sc.hint_tag && h > sc.next) ? h : sc.next;
}

// store entry e for sweep `target` (the caller has raised the epoch and owns the push)
__device__ __forceinline__ void push_entry(const Sched& sc, unsigned target, i64 e) {
    if (target == sc.next) {
        const i64 pos = sc.base + warp_agg_reserve(&sc.ctr->tail3[sc.tslot], 0u);
        if (pos - sc.begin < sc.ring.cap) sc.ring.slots[pos & (sc.ring.cap - 1)] = e;
        else *reinterpret_cast<volatile i64*>(&sc.ctr->overflow[sc.ovf_slot]) = 1;
    } else {
        const i64 pos = warp_agg_reserve(&sc.ctr->fill[target], target);
        sc.park[sc.ctr->offs[target] + pos] = e;
        if (pos == 0 || (i64)target > *reinterpret_cast<volatile i64*>(&sc.ctr->last_bucket))
            atomicMax(reinterpret_cast<u64*>(&sc.ctr->last_bucket), (u64)target);
    }
}

// queue voxel n (packed entry e) unless it already is queued for that sweep or a later one
__device__ __forceinline__ void enqueue(const Sched& sc, i64 n, i64 e) {
    const unsigned target = target_sweep(sc, sc.aux[n].lvl);
    if (atomicMax(&sc.aux[n].epoch, target) >= target) return;
    push_entry(sc, target, e);
}

// far direct neighbours of a voxel that is frozen (or whose value entered the band) become
// candidates and are queued for their first evaluation
template <int NDIM>
__device__ __forceinline__ void expose_neighbours(i64 e, i64 l, const int (&coord)[3],
                                                  const FmmGeom& g, u8* mask, i64* list, i64 cap,
                                                  const Sched& sc) {
    const int ext[3] = {g.n0, g.n1, g.n2};
    const i64 st[3] = {(i64)g.n1 * g.n2, (i64)g.n2, 1};
    const i64 est[3] = {(i64)1 << g.sh0, (i64)1 << g.sh1, 1};
#pragma unroll
    for (int ax = 3 - NDIM; ax < 3; ++ax) {
#pragma unroll
        for (int j = -1; j < 2; j += 2) {
            const int cc = coord[ax] + j;
            if (cc < 0 || cc >= ext[ax]) continue;
            if (ax == 0 && (cc < g.own_lo || cc >= g.own_hi)) continue;       // a halo copy
            const i64 n = l + j * st[ax];
            if (__ldcg(mask + n) == 0 && claim(mask, n, ST_CAND)) {
                const i64 pos = (i64)atomicAdd(reinterpret_cast<u64*>(&sc.ctr->n_cand), (u64)1);
                list[cap - 1 - pos] = e + j * est[ax];
                enqueue(sc, n, e + j * est[ax]);
            }
        }
    }
}

template <int NDIM>
__global__ void __launch_bounds__(256)
fmm_seed_kernel(u8* mask, Aux* aux, i64* list, i64 cap, FmmGeom g, FmmCounters* ctr, Ring ring,
                i64* park) {
    const i64 ni = ctr->n_init;
    const unsigned tag = ctr->use_hints ? hint_tag_of(ctr->gen - 1) : 0xffffffffu;
    const Sched sc{aux, ring, park, ctr, 0, 0, 1u, 1, tag, 2};
    for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < ni;
         t += (i64)gridDim.x * blockDim.x) {
        int coord[3];
        const i64 e = list[t];
        const i64 l = decode(e, g, coord);
        expose_neighbours<NDIM>(e, l, coord, g, mask, list, cap, sc);
    }
}
// [host-sched:end]

// [host-test:begin] (tests/test_fmm_device_code_host.py compiles the marked regions with g++)
// ---- the marcher's upwind update on a local stencil ------------------------------------------
// Stencil slots are numbered by their RANK in flat-index order around x (rank 6 = x itself):
//   0: -2 along axis 0   1: -1 axis 0   2: -2 axis 1   3: -1 axis 1   4: -2 axis 2   5: -1 axis 2
//   7: +1 axis 2   8: +2 axis 2   9: +1 axis 1   10: +2 axis 1   11: +1 axis 0   12: +2 axis 0
// (whenever two slots both exist their flat indices are ordered like their ranks).  Direct
// neighbours have odd ranks; the voxel in between a two-away slot r and x is r+1 (r < 6) or r-1.
__device__ __forceinline__ constexpr int slot_rank(int ax, int d, int which) {
    return d == 0 ? (2 * ax + (which == 1 ? 1 : 0)) : (7 + 2 * (2 - ax) + (which == 1 ? 0 : 1));
}

// Per-axis contributors of the upwind update (oracle/lsml_oracle.c:fmm_update): the frozen direct
// neighbour with the smaller magnitude supplies v1; the voxel two cells away in the same
// direction supplies v2 when it is frozen and passes scikit-fmm's second-order switch.
struct AxisState {
    double v1, v2;
    float k1, k2;        // float32 ordering magnitudes of v1 / v2
    bool active, second;
    int dir;             // direction (0 = -1, 1 = +1) v1 was taken from
};

// scikit-fmm's switch `(d2 <= v1 && v1 >= 0) || (d2 >= v1 && v1 <= 0)` on rounded magnitudes
// (oracle/lsml_oracle.c:second_order_ok)
__device__ __forceinline__ bool second_order_ok(double d2, float q2, double v1, float k1) {
    return v1 == 0.0 || d2 * v1 < 0 || q2 <= k1;
}

// Contributor rule (oracle/lsml_oracle.c:across): a second neighbour of OPPOSITE
// sign lies across the interface -- its magnitude says nothing about causality -- and is exempt
// from the causality test and from eviction.  -DLSML_FMM_RULE_LEGACY (with the oracle's
// LSML_ORACLE_RULE_R2=0) restores the rule used before, for the regression test of the difference.
__device__ __forceinline__ bool across(const AxisState& a) {
#ifndef LSML_FMM_RULE_LEGACY
    return a.v2 * a.v1 < 0;
#else
    (void)a;
    return false;
#endif
}

// The quadratic of oracle/lsml_oracle.c:fmm_solve on the contributors in `ax` -- the same
// expressions in the same order, including the causality rule: a root that is not larger in
// magnitude than a value it was computed from (or a non-positive discriminant) removes the
// contributor with the largest magnitude and solves again.  Returns 0.0 without contributors.
template <int NDIM>
__device__ __forceinline__ double solve_quadratic(const AxisState (&in)[3], const FmmGeom& g,
                                                  bool phi_pos) {
    const double aa = 9.0 / 4.0, one_third = 1.0 / 3.0;
    bool active[3], second[3];
    int nactive = 0;
#pragma unroll
    for (int ax = 3 - NDIM; ax < 3; ++ax) {
        active[ax] = in[ax].active; second[ax] = in[ax].second;
        nactive += active[ax] ? 1 : 0;
    }
#pragma unroll 1
    while (nactive > 0) {
        double a = 0, b = 0, c = 0;
#pragma unroll
        for (int ax = 3 - NDIM; ax < 3; ++ax) {
            if (!active[ax]) continue;
            if (second[ax]) {
                const double tp = one_third * (4 * in[ax].v1 - in[ax].v2);
                a += g.idx2[ax] * aa;
                b -= g.idx2[ax] * 2 * aa * tp;
                c += g.idx2[ax] * aa * (tp * tp);
            } else {
                a += g.idx2[ax];
                b -= g.idx2[ax] * 2 * in[ax].v1;
                c += g.idx2[ax] * (in[ax].v1 * in[ax].v1);
            }
        }
        c -= 1;
        const double det = b * b - 4 * a * c;
        if (det > 0) {
            const double sq = sqrt(det);
            const double r = (phi_pos ? (-b + sq) : (-b - sq)) / 2.0 / a;
            const float qr = qmag(r);
            bool causal = true;
#pragma unroll
            for (int ax = 3 - NDIM; ax < 3; ++ax) {
                if (!active[ax]) continue;
                if (qr <= in[ax].k1) causal = false;
                if (second[ax] && !across(in[ax]) && qr <= in[ax].k2) causal = false;
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
            if (second[ax] && !across(in[ax]) && in[ax].k2 > wv) { wv = in[ax].k2; worst = ax; worst_is_v2 = true; }
            if (in[ax].k1 > wv) { wv = in[ax].k1; worst = ax; worst_is_v2 = false; }
        }
#pragma unroll
        for (int ax = 3 - NDIM; ax < 3; ++ax)
            if (ax == worst) {
                if (worst_is_v2) second[ax] = false;
                else active[ax] = false;
            }
        if (!worst_is_v2) nactive--;
    }
    return 0.0;
}

// V(x): replay of the marcher around x (see the header).  Returns +inf when x never gets a value.
//
// The stencil voxels that will be frozen before x ("events") are applied in key order.  The
// marcher re-evaluates x's update after every one of them from ALL frozen stencil voxels; here
// the per-axis contributors are maintained incrementally and the quadratic is only solved when
// an event actually changed them (an unchanged contributor set gives the identical value):
//   * a direct neighbour freezing on an axis that already has a frozen direct neighbour only
//     replaces it when fmm_update's own loop would prefer it (possible against an INITIAL frozen
//     voxel; among marched voxels events arrive in increasing key order and the earlier stays);
//   * a voxel two cells away matters only if the voxel in between supplies v1 of that axis.
template <int NDIM>
__device__ double replay(i64 l, const int (&coord)[3], const double* __restrict__ u,
                         const double* T, const u8* mask, const FmmGeom& g, double band,
                         float& reach, unsigned& level) {
    const int ext[3] = {g.n0, g.n1, g.n2};
    const i64 st[3] = {(i64)g.n1 * g.n2, (i64)g.n2, 1};
    double tv[13];                    // by slot rank; dynamically indexed => local memory (L1)
    unsigned char lv[13];             // dependency depth of the slot's voxel
    float qr[13];                     // the same magnitudes again in registers for the selection
    unsigned F = 0u;                  // frozen now
    unsigned E = 0u;                  // pending event (marched voxel with a value inside the band)
    const double inf = pos_inf();
    tv[6] = inf; qr[6] = 0.f; lv[6] = 0;
    // ALL stencil loads are issued before any of them is used: 24 independent L2 reads, one round
    // trip.  (Written as "if in range { load; use }" per slot, the loads sit in 12 separate
    // conditional blocks and each block waits for its own round trip: 12 in a row, a quarter of
    // the replay's latency in round 2's profile.)  An out-of-range slot reads x itself.
    double tl[13];
    u8 sl[13];
    unsigned okm = 0u;
#pragma unroll
    for (int ax = 3 - NDIM; ax < 3; ++ax) {
#pragma unroll
        for (int d = 0; d < 2; ++d) {
            const int j = d ? 1 : -1;
#pragma unroll
            for (int which = 1; which <= 2; ++which) {
                const int cc = coord[ax] + which * j;
                const bool ok = cc >= 0 && cc < ext[ax];
                const i64 n = ok ? l + which * j * st[ax] : l;
                const int r = slot_rank(ax, d, which);
                tl[r] = __ldcg(T + n);            // L2: other CTAs update T concurrently
                sl[r] = __ldcg(mask + n);
                okm |= ok ? 1u << r : 0u;
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 13; ++r) {
        if (r == 6) continue;
        if (NDIM < 3 && (r < 2 || r > 10)) { tv[r] = inf; qr[r] = 0.f; lv[r] = 0; continue; }
        if (NDIM < 2 && (r < 4 || r > 8)) { tv[r] = inf; qr[r] = 0.f; lv[r] = 0; continue; }
        const bool ok = (okm >> r) & 1u;
        const double t = tl[r];
        const u8 s = sl[r];
        tv[r] = ok ? t : inf;
        qr[r] = ok ? qmag(t) : 0.f;
        lv[r] = ok ? (unsigned char)(s >> 2) : (unsigned char)0;
        F |= (ok && s == ST_INIT) ? 1u << r : 0u;
        E |= (ok && s != ST_INIT && (s & ST_CAND) && fabs(t) <= band) ? 1u << r : 0u;
    }
    float ql[13];                     // dynamically indexed copy
#pragma unroll
    for (int r = 0; r < 13; ++r) ql[r] = qr[r];
    const bool phi_pos = __ldg(u + l) > DBL_EPSILON;

    // contributors from the initial frozen set (oracle/lsml_oracle.c:fmm_update's loop)
    AxisState axs[3];
    bool any_direct = false;
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        axs[ax].active = false; axs[ax].second = false; axs[ax].dir = 0;
        axs[ax].v1 = 0.0; axs[ax].v2 = 0.0; axs[ax].k1 = 0.f; axs[ax].k2 = 0.f;
    }
#pragma unroll
    for (int ax = 3 - NDIM; ax < 3; ++ax) {
#pragma unroll
        for (int d = 0; d < 2; ++d) {
            const int r1 = slot_rank(ax, d, 1), r2 = slot_rank(ax, d, 2);
            if (((F >> r1) & 1u) && (!axs[ax].active || qr[r1] < axs[ax].k1)) {
                axs[ax].active = true; axs[ax].dir = d;
                axs[ax].v1 = tv[r1]; axs[ax].k1 = qr[r1];
                axs[ax].v2 = tv[r2]; axs[ax].k2 = qr[r2];
                axs[ax].second = ((F >> r2) & 1u) &&
                                 second_order_ok(tv[r2], qr[r2], tv[r1], qr[r1]);
                any_direct = true;
            }
        }
    }
    bool narrow = false;
    double t = inf;
    unsigned self_q = 0xffffffffu;
    if (any_direct) {
        const double d0 = solve_quadratic<NDIM>(axs, g, phi_pos);
        if (d0 != 0.0) { t = d0; narrow = true; self_q = __float_as_uint(qmag(t)); }
    }
    bool stale = false;               // contributors changed since the last solve
    float last_q = 0.f;               // magnitude of the last consumed event
    unsigned deepest = 0u;            // largest depth among the consumed events
#pragma unroll 1
    while (E != 0u) {
        // next event: smallest (magnitude, rank) among the pending slots
        unsigned best_q = 0xffffffffu, best_r = 15u;
#pragma unroll
        for (int r = 0; r < 13; ++r) {
            if (r == 6) continue;
            if (NDIM < 3 && (r < 2 || r > 10)) continue;
            if (NDIM < 2 && (r < 4 || r > 8)) continue;
            const unsigned q = __float_as_uint(qr[r]);
            if (((E >> r) & 1u) && q < best_q) { best_q = q; best_r = (unsigned)r; }
        }
        // x is frozen before that voxel
        if (narrow && (self_q < best_q || (self_q == best_q && 6u < best_r))) break;
        const unsigned r = best_r;
        E &= ~(1u << r);
        F |= 1u << r;
        const bool lower = r < 6u;
        const unsigned rr = lower ? r : r - 7u;
        const int eax = lower ? (int)(rr >> 1) : 2 - (int)(rr >> 1);
        const int ed = lower ? 0 : 1;
        const bool direct = (r & 1u) != 0u;
        const double te = tv[r];
        const float qe = ql[r];
        last_q = qe;
        deepest = max(deepest, (unsigned)lv[r]);
        bool changed = false;
        if (direct) {
            const unsigned r2 = lower ? r - 1u : r + 1u;        // two cells away, same direction
            const double t2 = tv[r2];
            const float q2 = ql[r2];
            const bool fr2 = (F >> r2) & 1u;
#pragma unroll
            for (int ax = 3 - NDIM; ax < 3; ++ax)
                // takes over v1 when the axis has none yet, or when it beats the frozen direct
                // neighbour on the other side exactly as fmm_update's loop decides (-1 side
                // first, +1 side only if strictly smaller); among marched voxels the earlier
                // event always stays, an initial frozen voxel can be overtaken
                if (ax == eax && (!axs[ax].active ||
                                  (ed == 1 ? qe < axs[ax].k1 : qe <= axs[ax].k1))) {
                    axs[ax].active = true; axs[ax].dir = ed;
                    axs[ax].v1 = te; axs[ax].k1 = qe; axs[ax].v2 = t2; axs[ax].k2 = q2;
                    axs[ax].second = fr2 && second_order_ok(t2, q2, te, qe);
                    changed = true;
                }
        } else {
#pragma unroll
            for (int ax = 3 - NDIM; ax < 3; ++ax)
                if (ax == eax && axs[ax].active && axs[ax].dir == ed && !axs[ax].second &&
                    second_order_ok(te, qe, axs[ax].v1, axs[ax].k1)) {
                    axs[ax].v2 = te; axs[ax].k2 = qe; axs[ax].second = true;
                    changed = true;
                }
        }
        // the marcher updates x when a direct neighbour freezes, and when a voxel two cells away
        // freezes behind a frozen one while x is already narrow
        const unsigned rmid = lower ? r + 1u : r - 1u;
        const bool evaluates = direct || (((F >> rmid) & 1u) && narrow);
        if (evaluates) {
            if (changed || stale) {
                const double dn = solve_quadratic<NDIM>(axs, g, phi_pos);
                if (dn != 0.0) { t = dn; narrow = true; self_q = __float_as_uint(qmag(t)); }
                stale = false;
            }
        } else if (changed) {
            stale = true;
        }
    }
    reach = narrow ? fmaxf(last_q, qmag(t)) : __int_as_float(0x7f800000);
    level = min(deepest + 1u, MAX_LEVEL);
    return narrow ? t : inf;
}
// [host-test:end]

// [host-sched:begin] (tests/test_fmm_scheduler_host.py: the scheduler's device functions on one host thread)
constexpr int FMM_MAX_SWEEPS = 1 << 16;   // grid-wide sweeps per launch; ~1 s at worst, then
                                          // `converged` = 0 is reported (never seen)
// Sweeps per REBUILD, counting the single-CTA ones (a grid-wide step can hold 4096 of those, so
// FMM_MAX_SWEEPS alone does not bound the time).  Far above any dependency depth (a full 1024^3
// distance map needs ~3 000); a march that gets here would be cycling (never observed; the
// contributor rule of DESIGN.md 4.6 is what rules it out) and ends with `converged` = 0 after a
// fraction of a second instead of never.
constexpr unsigned FMM_SWEEP_LIMIT = 1u << 16;
// (the same 12 warps per SM as 2 CTAs of 192 or 1 CTA of 384 threads -- fewer participants in the
// grid barrier -- measured 992 / 1131 and 1031 / 1144 us per 256^3 rebuild against 1004 / 1132)
#ifndef FMM_THREADS_PER_CTA
#define FMM_THREADS_PER_CTA 128
#endif
constexpr int FMM_THREADS = FMM_THREADS_PER_CTA;
constexpr int FMM_WARPS = FMM_THREADS / 32;
// CTAs per SM the kernel is compiled for.  3 (166 registers per thread: no local-memory traffic
// for the replay's state) beats 4 (128 registers) on one volume -- 1011 / 1113 us against 1036 /
// 1195 us per 256^3 rebuild: a sweep is bound by the latency of ONE lane's replay, not by the
// number of warps -- and ties on batches (1583 vs 1587 ms per C4 engine batch); 5 (96) loses.
#ifndef FMM_MIN_BLOCKS
#define FMM_MIN_BLOCKS 3
#endif
// Dirty sets this small are swept by ONE CTA with block barriers (two entries per warp at most:
// a sweep then costs one replay + publish, ~8 us, instead of the grid-wide floor of 13.5 us).
// Up to 256 entries, as first built, the tail COST time: 32 unrelated entries share a warp as
// divergent lanes, ~20 us per sweep -- 1072 / 1252 / 1311 us per 256^3 rebuild against 1036 /
// 1195 / 1203 us with 8 (and 1026 / 1166 / 1195 us with the tail switched off altogether).
#ifndef FMM_TAIL_MAX_ENTRIES
#define FMM_TAIL_MAX_ENTRIES 8
#endif
constexpr int FMM_TAIL_MAX = FMM_TAIL_MAX_ENTRIES;

// x (entry e, index l) changed from magnitude q_old to q_new during sweep `cur`: queue, for the
// next sweep, every candidate whose stencil contains x and whose result can depend on it.
//   * a candidate that is itself in the work set of THIS sweep (epoch == cur) may have read
//     either value and its `reach` is in flux: queued unconditionally;
//   * any other candidate's `reach` is stable (written before the last barrier) and describes
//     a replay that saw x's old value: queued only if x is, or was, within its reach.
// When x's value entered the band (`expose`), its far direct neighbours become candidates and
// are queued for their first evaluation.
// Written as straight-line predicated phases (state bytes -> records -> claims/atomicMax -> one
// reservation per list) so that the dozen L2 round trips of each phase overlap instead of
// forming one long dependent chain per neighbour.
template <int NDIM>
__device__ __forceinline__ void publish_change(i64 e, i64 l, const int (&coord)[3], float q_min,
                                               bool expose, const FmmGeom& g, u8* mask, i64* list,
                                               i64 cap, const Sched& sc) {
    const int ext[3] = {g.n0, g.n1, g.n2};
    const i64 st[3] = {(i64)g.n1 * g.n2, (i64)g.n2, 1};
    const i64 est[3] = {(i64)1 << g.sh0, (i64)1 << g.sh1, 1};
    const unsigned cur = sc.next - 1u;
    constexpr int NS = 4 * NDIM;
    const int offs[4] = {-2, -1, 1, 2};
    // phase 1: state bytes of the stencil.  Every phase issues ALL its loads / atomics before it
    // looks at any result (an out-of-range slot reads x itself), so that a phase costs one L2
    // round trip and not one per neighbour.
    unsigned valid = 0u;
    u8 s[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        const int ax = 3 - NDIM + k / 4, j = offs[k % 4];
        const int cc = coord[ax] + j;
        const bool okk = cc >= 0 && cc < ext[ax] && (ax != 0 || (cc >= g.own_lo && cc < g.own_hi));
        valid |= okk ? 1u << k : 0u;              // (halo copies are neither queued nor claimed)
        s[k] = __ldcg(mask + (okk ? l + j * st[ax] : l));
    }
    // ... and, in the same round trip, the scheduling records of the stencil
    uint2 a[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        const int ax = 3 - NDIM + k / 4, j = offs[k % 4];
        a[k] = __ldcg(reinterpret_cast<const uint2*>(sc.aux + (((valid >> k) & 1u) ? l + j * st[ax] : l)));
    }
#pragma unroll
    for (int k = 0; k < NS; ++k)
        if (!((valid >> k) & 1u)) s[k] = 0xff;
    // phase 2: claims of far direct neighbours (only when x's value has just entered the band)
    unsigned oldw[NS];
    unsigned claims = 0u, won = 0u;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        const int ax = 3 - NDIM + k / 4, j = offs[k % 4];
        const i64 n = ((valid >> k) & 1u) ? l + j * st[ax] : l;
        oldw[k] = 0xffffffffu;
        if (expose && (j == 1 || j == -1) && s[k] == 0) {        // claim(): state 0 -> candidate
            claims |= 1u << k;
            oldw[k] = atomicOr(reinterpret_cast<unsigned*>(mask + (n & ~(i64)3)),
                               (unsigned)ST_CAND << (8u * (unsigned)(n & 3)));
        }
    }
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        const int ax = 3 - NDIM + k / 4, j = offs[k % 4];
        const i64 n = l + j * st[ax];
        const bool claimed = (claims >> k) & 1u;
        if (claimed && ((oldw[k] >> (8u * (unsigned)(n & 3))) & 0xffu) == 0u) won |= 1u << k;
        if (!((s[k] & 3) == ST_CAND || claimed)) a[k] = make_uint2(0xffffffffu, 0u);
    }
    // phase 3: queue (atomicMax on the epoch; the thread that raises it owns the push)
    unsigned prev[NS], target[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        const int ax = 3 - NDIM + k / 4, j = offs[k % 4];
        const i64 n = l + j * st[ax];
        target[k] = target_sweep(sc, a[k].y >> 16);
        const bool want = ((won >> k) & 1u) ||
                          ((s[k] & 3) == ST_CAND && a[k].x < target[k] &&
                           (a[k].x == cur || q_min <= reach_from16(a[k].y & 0xffffu)));
        prev[k] = 0xffffffffu;
        if (want) prev[k] = atomicMax(&sc.aux[n].epoch, target[k]);
    }
    unsigned push = 0u, ringp = 0u;
#pragma unroll
    for (int k = 0; k < NS; ++k)
        if (prev[k] < target[k]) { push |= 1u << k; if (target[k] == sc.next) ringp |= 1u << k; }
    // phase 4: one reservation in the ring, one in the candidate list; parked entries go to
    // the bucket of their sweep
    // (aggregated over the lanes of the warp that are publishing together)
    const int nring = __popc(ringp), nwon = __popc(won);
    i64 pos, cpos;
    {
        // exclusive prefix sums of nring (<= 12: 4 bits) and nwon (<= 6: 3 bits) over the lanes
        // publishing together, one ballot per bit (cg::inclusive_scan on a coalesced group
        // loops over __fns: ~7 % of the kernel's instructions in round 2's profile)
        const unsigned active = __activemask();
        const unsigned lower = active & ((1u << (threadIdx.x & 31u)) - 1u);
        int er = 0, tr = 0, ew = 0, tw = 0;
#pragma unroll
        for (int bit = 0; bit < 4; ++bit) {
            const unsigned bal = __ballot_sync(active, (nring >> bit) & 1);
            er += __popc(bal & lower) << bit;
            tr += __popc(bal) << bit;
        }
#pragma unroll
        for (int bit = 0; bit < 3; ++bit) {
            const unsigned bal = __ballot_sync(active, (nwon >> bit) & 1);
            ew += __popc(bal & lower) << bit;
            tw += __popc(bal) << bit;
        }
        const int leader = __ffs(active) - 1;
        i64 br = 0, bw = 0;
        if ((int)(threadIdx.x & 31u) == leader) {
            if (tr) br = (i64)atomicAdd(reinterpret_cast<u64*>(&sc.ctr->tail3[sc.tslot]), (u64)tr);
            if (tw) bw = (i64)atomicAdd(reinterpret_cast<u64*>(&sc.ctr->n_cand), (u64)tw);
        }
        pos = sc.base + __shfl_sync(active, br, leader) + er;
        cpos = __shfl_sync(active, bw, leader) + ew;
    }
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        const int ax = 3 - NDIM + k / 4, j = offs[k % 4];
        const i64 en = e + j * est[ax];
        if ((ringp >> k) & 1u) {
            if (pos - sc.begin < sc.ring.cap) sc.ring.slots[pos & (sc.ring.cap - 1)] = en;
            else *reinterpret_cast<volatile i64*>(&sc.ctr->overflow[sc.ovf_slot]) = 1;
            ++pos;
        } else if ((push >> k) & 1u) {
            push_entry(sc, target[k], en);
        }
        if ((won >> k) & 1u) { list[cap - 1 - cpos] = en; ++cpos; }
    }
}

// One queued candidate: replay, publish a changed value.  T is updated in place; there are no
// fences -- a reader in the same sweep may see the old or the new value, and is queued again
// for the next sweep by the writer (see publish_change), after the barrier.
template <int NDIM>
__device__ __forceinline__ void process_candidate(i64 e, const double* __restrict__ u, double* T,
                                                  u8* mask, i64* list, i64 cap, const FmmGeom& g,
                                                  double band, const Sched& sc, unsigned gen_tag,
                                                  i64* hist, long long* clk = nullptr) {
    int coord[3];
    const i64 l = decode(e, g, coord);
    const double old = __ldcg(T + l);
    unsigned lvl = sc.aux[l].lvl;
    float reach;
    unsigned level;
    const double nw = replay<NDIM>(l, coord, u, T, mask, g, band, reach, level);
    const bool changed = __double_as_longlong(old) != __double_as_longlong(nw);
#ifdef FMM_PHASE_CLOCKS
    if (clk) clk[0] = clock64() + (changed ? 0 : 0);
#endif
    // depth: visible to the neighbours' replays through the state byte, and kept (tagged with
    // this rebuild) as the hint for the next rebuild
    mask[l] = (u8)((level << 2) | ST_CAND);
    // the hint is the SWEEP the voxel should be evaluated at next time: its depth, plus
    // (-DFMM_HINT_SLACK_SHIFT=n) depth >> n sweeps of slack that absorb small shifts of the
    // dependency structure upstream instead of cascading them into re-evaluations downstream
    // On BATCHES (thousands of items per rebuild: throughput-bound, a sweep costs what its entries
    // cost) the slack pays: -12 % replays for +20 % sweeps at depth / 4, C4 rebuild -4 %.  On one
    // volume (latency-bound: every sweep costs its 13.5 us floor) it costs +8..12 %, so it is
    // off there (fmm_rebuild: batch >= 16).
#if defined(FMM_HINT_SLACK_SHIFT)
    const unsigned hint = level + (level >> FMM_HINT_SLACK_SHIFT);
#elif defined(FMM_HINT_EARLY)
    // ... or FMM_HINT_EARLY sweeps BEFORE its depth: a voxel whose depth shrank since the last
    // rebuild is then still on time (a late evaluation delays everything downstream of it and
    // adds sweeps; an early one only costs a replay)
    const unsigned hint = level > FMM_HINT_EARLY ? level - FMM_HINT_EARLY : 1u;
#else
    const unsigned hint = sc.slack ? level + (level >> sc.slack) : level;
#endif
    if (lvl != (gen_tag | hint)) {
        if ((lvl & 0xff00u) == gen_tag) warp_agg_add(hist, lvl & 0xffu, -1);
        warp_agg_add(hist, hint, 1);
        lvl = gen_tag | hint;
    }
    // reach and hint share a 32-bit word that only this voxel's evaluator writes
    reinterpret_cast<unsigned*>(sc.aux + l)[1] = (unsigned)reach_up16(reach) | (lvl << 16);
#ifdef FMM_PHASE_CLOCKS
    if (clk) clk[1] = clock64();
#endif
    if (!changed) return;
    __stcg(T + l, nw);
    // a voxel whose value enters the band exposes its direct neighbours
    publish_change<NDIM>(e, l, coord, fminf(qmag(old), qmag(nw)),
                         fabs(nw) <= band && !(fabs(old) <= band), g, mask, list, cap, sc);
}

// [host-sched:end]

// Sweeps of the replay operator until no candidate is queued.  A candidate is queued when it is
// created and whenever a stencil voxel within its reach changes value; the thread that queues
// it appends it to the work ring (next sweep) or, when last rebuild's hint says its value
// settles later, to the parking bucket of that sweep.  The work set of sweep s is the ring
// window filled during sweep s-1 plus bucket s.  Large sets are swept by the whole grid with one
// grid barrier per sweep.  Small ones -- a few dozen voxels per sweep chasing dependency chains
// along the interface -- are swept by CTA 0 alone with block barriers.  If the ring overflows,
// the next sweep finds its work by scanning the candidate list for epoch == sweep.
template <int NDIM>
__global__ void __launch_bounds__(FMM_THREADS, FMM_MIN_BLOCKS)
fmm_march_kernel(const double* __restrict__ u, double* T, Aux* aux, u8* mask, i64* list, i64 cap,
                 FmmGeom g, double band, FmmCounters* ctr, Ring ring, i64* park, int finalize) {
    cg::grid_group grid = cg::this_grid();
    __shared__ i64 queue[FMM_WARPS][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const i64 tid = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const i64 nthreads = (i64)gridDim.x * blockDim.x;
    const i64 gwarp = tid >> 5, nwarps = nthreads >> 5;
    volatile i64* vtail = ctr->tail3;
    volatile i64* vovf = ctr->overflow;
    volatile i64* vpub = &ctr->pub_begin;
    volatile i64* vfill = ctr->fill;
    volatile i64* vlast = &ctr->last_bucket;
    const bool hints = ctr->use_hints != 0;
    const unsigned gen_tag = hint_tag_of(ctr->gen);
    const unsigned hint_tag = hints ? hint_tag_of(ctr->gen - 1) : 0xffffffffu;
    const unsigned slack = (unsigned)ctr->hint_slack;
    i64* const hist_cur = ctr->hist[ctr->gen & 1];
    // (a rebuild is one launch on a single GPU; a slab-decomposed rebuild launches again after
    // every halo exchange and continues where the previous launch stopped)
    unsigned sweep = (unsigned)ctr->resume_sweep;      // epoch of the work set being consumed
    unsigned rix = (unsigned)ctr->resume_rix;          // executed-sweep counter (rotation of tail3)
    int barriers = 0, tail_sweeps = 0;
    bool converged = false;
    i64 begin = ctr->resume_begin, end = begin + vtail[rix % 3u];   // queued by seed / requeue
    bool scan = vovf[2] != 0;          // ... which report overflow into slot 2
    i64 nreplay = 0;
    i64 nbk_pre = 0;
    bool pre = false;
    // Every CTA must have taken its snapshot of the start state before any CTA mutates it: a
    // small first work set sends CTA 0 straight into its single-CTA sweeps, which zero
    // tail3[rix % 3] from their second sweep on -- a CTA scheduled a few microseconds late
    // would then read end == begin, declare convergence and leave without ever reaching the
    // grid barrier the others wait at.  All later loop-exit tests only read values published
    // before a barrier (or monotone ones: fill[sweep] is final before sweep `sweep` starts,
    // last_bucket only grows and is compared while nothing is being processed).
    grid.sync();
    while (barriers < FMM_MAX_SWEEPS && sweep < FMM_SWEEP_LIMIT) {
#ifdef FMM_PHASE_CLOCKS
        const long long loop_top = clock64();
#endif
        // bucket of this sweep (already fetched with the ring window after the last barrier)
        i64 nbk = pre ? nbk_pre : ((hints && sweep < 256u) ? vfill[sweep] : 0);
        pre = false;
        if (end == begin && nbk == 0 && !scan) {
            if (!hints || (i64)sweep >= *vlast) { converged = true; break; }
            ++sweep;                    // nothing queued for this sweep, something parked later
            continue;
        }
        const int slot = barriers % 3;
        if (tid == 0) {
            vovf[(barriers + 1) % 3] = 0;
            vtail[(rix + 2u) % 3u] = 0;
            volatile i64* live = ctr->live;
            live[0] = (i64)sweep; live[1] = (i64)rix; live[2] = begin; live[3] = end;
            if (barriers < 32) {
                ctr->sizes[barriers] = scan ? -1 : end - begin + nbk;
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                ctr->stamps[barriers] = (i64)now;
            }
        }
        if (!scan && end - begin + nbk <= FMM_TAIL_MAX) {
            // ---- a run of small sweeps on CTA 0; everybody else waits at the barrier
            if (blockIdx.x == 0) {
                i64 b = begin, e = end;
                int inner = 0;
                while (inner < 4096 && sweep < FMM_SWEEP_LIMIT) {
                    nbk = (hints && sweep < 256u) ? vfill[sweep] : 0;
                    const i64 work = e - b + nbk;
                    if (work == 0) {
                        if (!hints || (i64)sweep >= *vlast) break;
                        ++sweep;
                        continue;
                    }
                    if (work > FMM_TAIL_MAX) break;
                    if (threadIdx.x == 0 && inner > 0) vtail[(rix + 2u) % 3u] = 0;
                    const Sched sc{aux, ring, park, ctr, b, e, sweep + 1u, (int)((rix + 1u) % 3u),
                                   hint_tag, slot, slack};
                    const i64 pbase = nbk ? ctr->offs[sweep] : 0;
                    // (entry k to warp k % 4, lane k / 4: the few entries of a tail sweep each get
                    // a warp of their own instead of sharing one as divergent lanes)
                    for (i64 k = (i64)lane * FMM_WARPS + warp; k < work; k += FMM_THREADS) {
                        const i64 ent = k < e - b ? ring.slots[(b + k) & (ring.cap - 1)]
                                                  : park[pbase + k - (e - b)];
                        process_candidate<NDIM>(ent, u, T, mask, list, cap, g, band, sc, gen_tag, hist_cur);
                        ++nreplay;
                    }
                    __threadfence();
                    __syncthreads();
                    b = e;
                    e = b + vtail[(rix + 1u) % 3u];
                    ++sweep;
                    ++rix;
                    ++inner;
                    const bool dropped = vovf[slot] != 0;      // (never seen: the ring is >= 64 Ki)
                    __syncthreads();
                    if (dropped) break;
                }
                if (threadIdx.x == 0) {
                    vpub[0] = b; vpub[1] = e; vpub[2] = (i64)sweep; vpub[3] = (i64)rix;
                    ctr->tail_sweeps += inner;
                }
                tail_sweeps += inner;
            }
            __threadfence();
            grid.sync();
            begin = vpub[0];
            end = vpub[1];
            sweep = (unsigned)vpub[2];
            rix = (unsigned)vpub[3];
            scan = vovf[slot] != 0;
            ++barriers;
            continue;
        }
        if (!scan) {
            const Sched sc{aux, ring, park, ctr, begin, end, sweep + 1u, (int)((rix + 1u) % 3u),
                           hint_tag, slot, slack};
            const i64 work = end - begin + nbk;
            const i64 pbase = nbk ? ctr->offs[sweep] : 0;
            // The work set is dealt to the warps in CHUNKS of G consecutive entries (entries
            // pushed together belong to neighbouring voxels: similar stencils, similar replay
            // paths, overlapping cache lines), chunk c to warp c % nwarps; G is the smallest power
            // of two with which no warp gets a second, unrelated chunk (G = 1: a small set is
            // spread one entry per warp -- a lone lane runs its own path, not the union of 32
            // divergent ones; G = 32: a large one fills whole warps with consecutive entries).
            // Measured against "largest G that still gives every warp work" (-DFMM_CHUNK_FLOOR):
            // 1191 vs 1234 us per 256^3 rebuild, equal on 2048-crop batches.
            // -DFMM_CHUNK=n fixes G (probes, tools/r2_fmm_variants.sh).
            int G = 1;
#ifdef FMM_CHUNK
            G = FMM_CHUNK;
#elif defined(FMM_CHUNK_FLOOR)
            while (G < 32 && (i64)(2 * G) * nwarps <= work) G <<= 1;   // every warp gets work
#else
            while (G < 32 && (i64)G * nwarps < work) G <<= 1;          // at most one chunk per warp
#endif
            const int per_warp = 32 / G;             // chunks a warp holds per pass
#ifdef FMM_PHASE_CLOCKS
            const bool sampled = tid == 32 && work > 32 && work <= nwarps;
            long long ck[6] = {0, 0, 0, 0, 0, 0};
#endif
            for (i64 round = lane / G;; round += per_warp) {
                const i64 k = (round * nwarps + gwarp) * G + (lane & (G - 1));
                if (k >= work) break;
                const i64 ent = k < end - begin ? ring.slots[(begin + k) & (ring.cap - 1)]
                                                : park[pbase + k - (end - begin)];
#ifdef FMM_PHASE_CLOCKS
                if (sampled) {
                    ck[0] = clock64() + (ent & 0);   // (after the entry has arrived)
                    process_candidate<NDIM>(ent, u, T, mask, list, cap, g, band, sc, gen_tag, hist_cur, ck + 1);
                    ck[3] = clock64();
                } else
#endif
                process_candidate<NDIM>(ent, u, T, mask, list, cap, g, band, sc, gen_tag, hist_cur);
                ++nreplay;
            }
#ifdef FMM_PHASE_CLOCKS
            if (sampled) {
                __threadfence();
                ck[4] = clock64();
                grid.sync();
                ck[5] = clock64();
                ctr->phase[0] += ck[0] - loop_top;
                ctr->phase[1] += ck[1] - ck[0];
                ctr->phase[2] += ck[2] - ck[1];
                ctr->phase[3] += ck[3] - ck[2];
                ctr->phase[4] += ck[4] - ck[3];
                ctr->phase[5] += ck[5] - ck[4];
                ctr->phase[7] += 1;
                begin = end;
                end = begin + vtail[(rix + 1u) % 3u];
                scan = vovf[slot] != 0;
                ++sweep; ++rix; ++barriers;
                continue;
            }
#endif
        } else {
            // ---- fallback: scan the candidate list; a warp collects its queued entries in
            // shared memory and evaluates them 32 at a time
            const Sched sc{aux, ring, park, ctr, end, end, sweep + 1u, (int)((rix + 1u) % 3u),
                           hint_tag, slot, slack};
            const i64 nc = *reinterpret_cast<volatile i64*>(&ctr->n_cand);
            int qn = 0;
            for (i64 base = gwarp * 32; base < nc; base += nwarps * 32) {
                const i64 t = base + lane;
                i64 e = -1;
                bool queued = false;
                if (t < nc) {
                    e = list[cap - 1 - t];
                    int coord[3];
                    queued = __ldcg(&aux[decode(e, g, coord)].epoch) == sweep;
                }
                const unsigned bal = __ballot_sync(0xffffffffu, queued);
                if (queued) queue[warp][qn + __popc(bal & ((1u << lane) - 1u))] = e;
                qn += __popc(bal);
                __syncwarp();
                if (qn >= 32) {
                    const i64 e2 = queue[warp][qn - 32 + lane];
                    qn -= 32;
                    __syncwarp();
                    process_candidate<NDIM>(e2, u, T, mask, list, cap, g, band, sc, gen_tag, hist_cur);
                    ++nreplay;
                }
            }
            if (lane < qn) {
                process_candidate<NDIM>(queue[warp][lane], u, T, mask, list, cap, g, band, sc, gen_tag, hist_cur);
                ++nreplay;
            }
            __syncwarp();
        }
        // (grid.sync() alone -- CTA barrier + atom.add.release.gpu by one thread per CTA + acquire
        // polling, cooperative_groups/details/sync.h -- orders the sweep's writes as well and the
        // suite passes without the explicit fence; measured, dropping it saves nothing:
        // -DFMM_NO_EXTRA_FENCE 1100 / 1253 us vs 1096 / 1252 us per rebuild.  It stays.)
#ifndef FMM_NO_EXTRA_FENCE
        __threadfence();
#endif
        grid.sync();
        // ring entries behind `end` are complete unless this sweep overflowed, in which case the
        // next sweep scans instead.  The three words the next sweep's header needs are fetched
        // together: one L2 round trip.
        {
            const i64 queued = vtail[(rix + 1u) % 3u];
            const i64 dropped = vovf[slot];
            nbk_pre = (hints && sweep + 1u < 256u) ? vfill[sweep + 1u] : 0;   // final since the barrier
            pre = true;
            begin = end;
            end = begin + queued;
            scan = dropped != 0;
        }
        ++sweep;
        ++rix;
        ++barriers;
    }
    // final mask: candidates frozen by the marcher are those with |d| <= band
    const i64 nc = *reinterpret_cast<volatile i64*>(&ctr->n_cand);
    if (finalize) {
        for (i64 t = tid; t < nc; t += nthreads) {
            int coord[3];
            const i64 l = decode(list[cap - 1 - t], g, coord);
            if (fabs(__ldcg(T + l)) <= band) mask[l] = 1;
            else { mask[l] = 0; T[l] = pos_inf(); }
        }
    }
    if (nreplay) atomicAdd(reinterpret_cast<u64*>(&ctr->replays), (u64)nreplay);
    if (tid == 0) {
        ctr->sweeps = (i64)sweep - 1;
        ctr->converged = converged ? 1 : 0;
        if (!converged && finalize) ctr->failed_rebuilds += 1;
        ctr->barriers = barriers;
        ctr->prev_init = ctr->n_init;
        ctr->prev_cand = nc;
        ctr->resume_sweep = (i64)sweep;
        ctr->resume_rix = (i64)rix;
        ctr->resume_begin = begin;
    }
}

// ---- slab decomposition helpers ----------------------------------------------------------------
// After the halo planes were refreshed from the neighbour ranks: every owned candidate within two
// planes of a cut looks again, and far owned voxels next to a (possibly remote) frozen or in-band
// voxel become candidates -- exactly what the remote voxel's own thread does for neighbours it
// owns.  Entries are queued for the sweep the next march launch starts with.
template <int NDIM>
__global__ void __launch_bounds__(256)
fmm_requeue_kernel(const double* T, Aux* aux, u8* mask, i64* list, i64 cap, FmmGeom g,
                   double band, FmmCounters* ctr, Ring ring, i64* park,
                   const i64* __restrict__ sides_dev) {
    const int k = blockIdx.z * blockDim.x + threadIdx.x;
    const int cj = blockIdx.y * blockDim.y + threadIdx.y;
    if (k >= g.n2 || cj >= g.n1) return;
    // blockIdx.x enumerates the (up to) four boundary planes: lo, lo+1, hi-2, hi-1
    const int own = g.own_hi - g.own_lo;
    int ci;
    if (blockIdx.x == 0) ci = g.own_lo;
    else if (blockIdx.x == 1) ci = g.own_lo + 1;
    else if (blockIdx.x == 2) ci = g.own_hi - 2;
    else ci = g.own_hi - 1;
    if (ci < g.own_lo || ci >= g.own_hi) return;
    if (blockIdx.x >= 2 && ci < g.own_lo + 2) return;           // thin slab: planes already done
    if (own <= 0) return;
    // sides_dev[0] / [1] != 0: the halo below / above the owned planes changed in this exchange
    const bool near_lo = sides_dev[0] != 0 && g.own_lo > 0 && ci < g.own_lo + 2;
    const bool near_hi = sides_dev[1] != 0 && g.own_hi < g.n0 && ci >= g.own_hi - 2;
    if (!near_lo && !near_hi) return;
    const i64 l = ((i64)ci * g.n1 + cj) * g.n2 + k;
    const int coord[3] = {ci, cj, k};
    const i64 e = encode(0, ci, cj, k, g);
    const unsigned rix = (unsigned)ctr->resume_rix;
    const Sched sc{aux, ring, park, ctr, ctr->resume_begin, ctr->resume_begin,
                   (unsigned)ctr->resume_sweep, (int)(rix % 3u), 0xffffffffu, 2};
    const u8 st = __ldcg(mask + l);
    if (st & ST_CAND) { enqueue(sc, l, e); return; }
    if (st != 0) return;
    const int ext[3] = {g.n0, g.n1, g.n2};
    const i64 stv[3] = {(i64)g.n1 * g.n2, (i64)g.n2, 1};
    bool exposed = false;
#pragma unroll
    for (int ax = 3 - NDIM; ax < 3; ++ax)
#pragma unroll
        for (int j = -1; j < 2; j += 2) {
            const int cc = coord[ax] + j;
            if (cc < 0 || cc >= ext[ax]) continue;
            const i64 n = l + j * stv[ax];
            const u8 sn = __ldcg(mask + n);
            if (sn == ST_INIT || ((sn & ST_CAND) && fabs(__ldcg(T + n)) <= band)) exposed = true;
        }
    if (exposed && claim(mask, l, ST_CAND)) {
        const i64 pos = (i64)atomicAdd(reinterpret_cast<u64*>(&ctr->n_cand), (u64)1);
        list[cap - 1 - pos] = e;
        enqueue(sc, l, e);
    }
}

// the final mask of a slab-decomposed rebuild (the single-GPU march does this itself)
__global__ void __launch_bounds__(256)
fmm_finalize_kernel(double* T, u8* mask, const i64* __restrict__ list, i64 cap, FmmGeom g,
                    double band, const FmmCounters* ctr) {
    const i64 nc = ctr->n_cand;
    for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < nc;
         t += (i64)gridDim.x * blockDim.x) {
        int coord[3];
        const i64 l = decode(list[cap - 1 - t], g, coord);
        if (fabs(T[l]) <= band) mask[l] = 1;
        else { mask[l] = 0; T[l] = pos_inf(); }
    }
}

// number of entries waiting for the next march launch -> a device int64 (for the all-reduce)
__global__ void fmm_pending_kernel(const FmmCounters* ctr, i64* out) {
    *out = ctr->tail3[ctr->resume_rix % 3] + (ctr->overflow[2] ? 1 : 0);
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

__global__ void __launch_bounds__(256)
fill_aux_kernel(Aux* __restrict__ aux, i64 total) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (i64)gridDim.x * blockDim.x)
        aux[i] = Aux{0u, (unsigned short)REACH16_INF, (unsigned short)0};
}

// ---------------------------------------------------------------------------------- host side
// workspace layout: T (total f64) | aux (total x 8 B) | list (total i64) | ring (ring_cap i64)
// | parking array (ring_cap i64) | counters
struct FmmWorkspace {
    double* T;
    Aux* aux;
    i64* list;
    Ring ring;
    i64* park;
    FmmCounters* ctr;
};

static i64 ring_capacity(i64 total) {
    // a power of two (slot = position & (cap-1) ... kept as % for clarity) that holds every
    // dirty set seen in practice: a quarter of the grid, at least 64 Ki entries, at most total
    i64 want = total / 4 > 65536 ? total / 4 : 65536;
    i64 cap = 1;
    while (cap < want) cap <<= 1;
    return cap;
}

static FmmWorkspace carve(void* workspace, i64 total) {
    FmmWorkspace w;
    char* p = static_cast<char*>(workspace);
    w.T = reinterpret_cast<double*>(p); p += total * sizeof(double);
    w.aux = reinterpret_cast<Aux*>(p); p += total * sizeof(Aux);
    w.list = reinterpret_cast<i64*>(p); p += total * sizeof(i64);
    w.ring.slots = reinterpret_cast<i64*>(p); w.ring.cap = ring_capacity(total);
    p += w.ring.cap * sizeof(i64);
    w.park = reinterpret_cast<i64*>(p); p += w.ring.cap * sizeof(i64);
    w.ctr = reinterpret_cast<FmmCounters*>(p);
    return w;
}

static int g_march_blocks[4] = {0, 0, 0, 0};

template <int NDIM>
static int march_launch(const double* u, const FmmWorkspace& w, u8* mask, i64 cap, FmmGeom fg,
                        double band, int finalize, cudaStream_t s) {
    if (g_march_blocks[NDIM] == 0) {
        int per_sm = 0, dev = 0, sms = 0;
        LSML_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fmm_march_kernel<NDIM>,
                                                                FMM_THREADS, 0));
        LSML_CUDA(cudaGetDevice(&dev));
        LSML_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (per_sm < 1) { set_error("fmm_march_kernel does not fit on an SM"); return 1; }
        g_march_blocks[NDIM] = per_sm * sms;
    }
    double* T = w.T; Aux* aux = w.aux; i64* list = w.list; FmmCounters* ctr = w.ctr;
    Ring ring = w.ring;
    i64* park = w.park;
    void* args[] = {(void*)&u, (void*)&T, (void*)&aux, (void*)&mask, (void*)&list, (void*)&cap,
                    (void*)&fg, (void*)&band, (void*)&ctr, (void*)&ring, (void*)&park,
                    (void*)&finalize};
    LSML_CUDA(cudaLaunchCooperativeKernel((void*)fmm_march_kernel<NDIM>, dim3(g_march_blocks[NDIM]),
                                          dim3(FMM_THREADS), args, 0, s));
    count_launch();
    return 0;
}

i64 fmm_workspace_bytes(i64 total) {
    return total * (i64)(sizeof(double) + sizeof(Aux) + sizeof(i64)) +
           2 * ring_capacity(total) * (i64)sizeof(i64) +
           (i64)sizeof(FmmCounters);
}

// first use of a workspace: T = +inf everywhere, counters zero
int fmm_workspace_init(i64 total, void* workspace, cudaStream_t s) {
    const FmmWorkspace w = carve(workspace, total);
    fill_inf_kernel<<<LSML_SMS * 8, 256, 0, s>>>(w.T, total);
    LSML_LAUNCH_CHECK("fill_inf_kernel");
    fill_aux_kernel<<<LSML_SMS * 8, 256, 0, s>>>(w.aux, total);
    LSML_LAUNCH_CHECK("fill_aux_kernel");
    LSML_CUDA(cudaMemsetAsync(w.ctr, 0, sizeof(FmmCounters), s));
    return 0;
}

static void plane_launch_shape(const Geom& g, int planes, dim3& grid, dim3& block) {
    int bx = 32;
    while (bx < 256 && bx < g.n[2]) bx <<= 1;
    int by = 256 / bx;
    if (by > g.n[1]) { by = 1; while (by < g.n[1]) by <<= 1; }
    block = dim3(bx, by);
    grid = dim3((unsigned)(g.batch * planes), (unsigned)((g.n[1] + by - 1) / by),
                (unsigned)((g.n[2] + bx - 1) / bx));
}

// clear + prepare + initial frozen set + seeds
template <int NDIM>
static int begin_nd(const Geom& g, const FmmGeom& fg, const double* u, const u64* stats, u8* mask,
                    const FmmWorkspace& w, i64 total, const i64* prev_band_idx,
                    const i64* prev_n_band, const int (&dense)[4], cudaStream_t s) {
    static const bool no_hints = getenv("LSML_FMM_NO_HINTS") != nullptr;      // A/B switch (probes)
    fmm_clear_kernel<<<LSML_SMS * 8, 256, 0, s>>>(w.T, w.aux, mask, w.list, total, fg, w.ctr, total);
    LSML_LAUNCH_CHECK("fmm_clear_kernel");
    static const bool no_slack = getenv("LSML_FMM_NO_SLACK") != nullptr;      // A/B switch (probes)
    fmm_prepare_kernel<<<1, 256, 0, s>>>(w.ctr, (prev_band_idx != nullptr && !no_hints) ? 1 : 0,
                                         w.ring.cap, (g.batch >= 16 && !no_slack) ? 2 : 0);
    LSML_LAUNCH_CHECK("fmm_prepare_kernel");
    dim3 grid, block;
    if (prev_band_idx != nullptr) {
        fmm_init_band_kernel<NDIM><<<LSML_SMS * 8, 256, 0, s>>>(u, w.T, mask, w.list, fg, stats,
                                                              prev_band_idx, prev_n_band, w.ctr);
        LSML_LAUNCH_CHECK("fmm_init_band_kernel");
        // slab decomposition: planes whose interface status can change from the neighbour's side
        // (the ghost planes a rank evaluates redundantly, and its first/last owned plane) are
        // searched densely; claims make overlaps with the sparse search harmless
        for (int q = 0; q < 2; ++q) {
            const int a = dense[2 * q] < fg.own_lo ? fg.own_lo : dense[2 * q];
            const int b = dense[2 * q + 1] > fg.own_hi ? fg.own_hi : dense[2 * q + 1];
            if (b <= a) continue;
            plane_launch_shape(g, b - a, grid, block);
            fmm_init_kernel<NDIM><<<grid, block, 0, s>>>(u, w.T, mask, w.list, fg, stats, w.ctr, a, b);
            LSML_LAUNCH_CHECK("fmm_init_kernel");
        }
    } else {
        plane_launch_shape(g, fg.own_hi - fg.own_lo, grid, block);
        fmm_init_kernel<NDIM><<<grid, block, 0, s>>>(u, w.T, mask, w.list, fg, stats, w.ctr,
                                                    fg.own_lo, fg.own_hi);
        LSML_LAUNCH_CHECK("fmm_init_kernel");
    }
    fmm_seed_kernel<NDIM><<<LSML_SMS * 8, 256, 0, s>>>(mask, w.aux, w.list, total, fg, w.ctr, w.ring,
                                                     w.park);
    LSML_LAUNCH_CHECK("fmm_seed_kernel");
    return 0;
}

#define LSML_ND(call3, call2, call1) (g.ndim == 3 ? (call3) : g.ndim == 2 ? (call2) : (call1))

// Rebuild the band of every item from u.  mask must be all-zero on first use of the workspace
// and is thereafter maintained by this function.  band == 0 => whole grid (no narrow band).
// stats: the exact {u>0} statistics of the SAME u (global_stats), used for the reference's
// single-sign / all-zero edge cases.  dist_out may be null.
// prev_band_idx / prev_n_band (device; may be null): the compacted band of the PREVIOUS rebuild
// when u has since changed only on that band (the evolution loop) -- the initial frozen set is
// then found from the band instead of a dense pass over u.
int fmm_rebuild(const Geom& g, const double* u, double band, const u64* stats, u8* mask,
                double* dist_out, void* workspace, const i64* prev_band_idx,
                const i64* prev_n_band, cudaStream_t s) {
    const i64 total = g.voxels * g.batch;
    const FmmWorkspace w = carve(workspace, total);
    FmmGeom fg;
    if (make_fmm_geom(g, fg)) return 1;
    const double eff_band = band != 0.0 ? band : DBL_MAX;
    const int none[4] = {0, 0, 0, 0};
    int st = LSML_ND(begin_nd<3>(g, fg, u, stats, mask, w, total, prev_band_idx, prev_n_band, none, s),
                     begin_nd<2>(g, fg, u, stats, mask, w, total, prev_band_idx, prev_n_band, none, s),
                     begin_nd<1>(g, fg, u, stats, mask, w, total, prev_band_idx, prev_n_band, none, s));
    if (st) return st;
    st = LSML_ND(march_launch<3>(u, w, mask, total, fg, eff_band, 1, s),
                 march_launch<2>(u, w, mask, total, fg, eff_band, 1, s),
                 march_launch<1>(u, w, mask, total, fg, eff_band, 1, s));
    if (st) return st;
    if (dist_out != nullptr) {
        fmm_export_kernel<<<LSML_SMS * 8, 256, 0, s>>>(w.T, mask, dist_out, total);
        LSML_LAUNCH_CHECK("fmm_export_kernel");
    }
    return 0;
}

// ---- slab-decomposed rebuild: the same kernels, driven in phases by the host layer ----------
// (lsml_b200/core/slab.py): begin -> [march -> halo exchange of T/state -> requeue]* -> finish.
// The local grid is the rank's owned planes [own_lo, own_hi) plus halo planes; batch must be 1.
static int slab_geom(const Geom& g, int own_lo, int own_hi, FmmGeom& fg) {
    if (g.batch != 1) { set_error("slab decomposition: one volume per rank"); return 1; }
    if (make_fmm_geom(g, fg)) return 1;
    if (own_lo < 0 || own_hi > g.n[0] || own_lo >= own_hi) {
        set_error("slab decomposition: bad owned plane range [%d, %d)", own_lo, own_hi);
        return 1;
    }
    fg.own_lo = own_lo; fg.own_hi = own_hi;
    return 0;
}

// dense_planes: {a0, b0, a1, b1} -- plane ranges searched densely for the interface in addition
// to the previous band (ignored on a first rebuild, which searches every evaluated plane)
int fmm_slab_begin(const Geom& g, int own_lo, int own_hi, const double* u, u8* mask,
                   void* workspace, const i64* prev_band_idx, const i64* prev_n_band,
                   const int* dense_planes, cudaStream_t s) {
    const i64 total = g.voxels * g.batch;
    const FmmWorkspace w = carve(workspace, total);
    FmmGeom fg;
    if (slab_geom(g, own_lo, own_hi, fg)) return 1;
    const int dense[4] = {dense_planes[0], dense_planes[1], dense_planes[2], dense_planes[3]};
    return LSML_ND(begin_nd<3>(g, fg, u, nullptr, mask, w, total, prev_band_idx, prev_n_band, dense, s),
                   begin_nd<2>(g, fg, u, nullptr, mask, w, total, prev_band_idx, prev_n_band, dense, s),
                   begin_nd<1>(g, fg, u, nullptr, mask, w, total, prev_band_idx, prev_n_band, dense, s));
}

int fmm_slab_march(const Geom& g, int own_lo, int own_hi, const double* u, double band, u8* mask,
                   void* workspace, cudaStream_t s) {
    const i64 total = g.voxels * g.batch;
    const FmmWorkspace w = carve(workspace, total);
    FmmGeom fg;
    if (slab_geom(g, own_lo, own_hi, fg)) return 1;
    const double eff_band = band != 0.0 ? band : DBL_MAX;
    return LSML_ND(march_launch<3>(u, w, mask, total, fg, eff_band, 0, s),
                   march_launch<2>(u, w, mask, total, fg, eff_band, 0, s),
                   march_launch<1>(u, w, mask, total, fg, eff_band, 0, s));
}

int fmm_slab_requeue(const Geom& g, int own_lo, int own_hi, double band, const i64* sides_dev,
                     u8* mask, void* workspace, i64* pending_out, cudaStream_t s) {
    const i64 total = g.voxels * g.batch;
    const FmmWorkspace w = carve(workspace, total);
    FmmGeom fg;
    if (slab_geom(g, own_lo, own_hi, fg)) return 1;
    const double eff_band = band != 0.0 ? band : DBL_MAX;
    dim3 grid, block;
    plane_launch_shape(g, 4, grid, block);
    if (g.ndim == 3) fmm_requeue_kernel<3><<<grid, block, 0, s>>>(w.T, w.aux, mask, w.list, total, fg, eff_band, w.ctr, w.ring, w.park, sides_dev);
    else if (g.ndim == 2) fmm_requeue_kernel<2><<<grid, block, 0, s>>>(w.T, w.aux, mask, w.list, total, fg, eff_band, w.ctr, w.ring, w.park, sides_dev);
    else fmm_requeue_kernel<1><<<grid, block, 0, s>>>(w.T, w.aux, mask, w.list, total, fg, eff_band, w.ctr, w.ring, w.park, sides_dev);
    LSML_LAUNCH_CHECK("fmm_requeue_kernel");
    if (pending_out != nullptr) {
        fmm_pending_kernel<<<1, 1, 0, s>>>(w.ctr, pending_out);
        LSML_LAUNCH_CHECK("fmm_pending_kernel");
    }
    return 0;
}

int fmm_slab_finish(const Geom& g, int own_lo, int own_hi, double band, u8* mask, double* dist_out,
                    void* workspace, cudaStream_t s) {
    const i64 total = g.voxels * g.batch;
    const FmmWorkspace w = carve(workspace, total);
    FmmGeom fg;
    if (slab_geom(g, own_lo, own_hi, fg)) return 1;
    const double eff_band = band != 0.0 ? band : DBL_MAX;
    fmm_finalize_kernel<<<LSML_SMS * 8, 256, 0, s>>>(w.T, mask, w.list, total, fg, eff_band, w.ctr);
    LSML_LAUNCH_CHECK("fmm_finalize_kernel");
    if (dist_out != nullptr) {
        fmm_export_kernel<<<LSML_SMS * 8, 256, 0, s>>>(w.T, mask, dist_out, total);
        LSML_LAUNCH_CHECK("fmm_export_kernel");
    }
    return 0;
}

// device address of the counters {n_init, n_cand, tail_sweeps, sweeps, converged, prev_init,
// prev_cand, replays, ...} of the last rebuild
int fmm_counters_ptr(i64 total, void* workspace, const i64** out) {
    *out = reinterpret_cast<const i64*>(carve(workspace, total).ctr);
    return 0;
}

// byte offset of FmmCounters::sizes inside the counters (diagnostics)
i64 fmm_counters_sizes_offset() { return (i64)offsetof(FmmCounters, sizes); }
i64 fmm_counters_failed_offset() { return (i64)offsetof(FmmCounters, failed_rebuilds); }
// byte offset of the counters inside a workspace for `total` voxels
i64 fmm_counters_offset(i64 total) {
    return total * (i64)(sizeof(double) + sizeof(Aux) + sizeof(i64)) +
           2 * ring_capacity(total) * (i64)sizeof(i64);
}
i64 fmm_counters_stamps_offset() { return (i64)offsetof(FmmCounters, stamps); }
i64 fmm_counters_phase_offset() { return (i64)offsetof(FmmCounters, phase); }

}  // namespace lsml
