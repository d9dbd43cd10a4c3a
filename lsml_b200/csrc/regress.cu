// regress.cu -- inference of the host-trained regressor on the band voxels
// (call sites lsml/core/model.py:388, fit_job_handler.py:504).
//
// The models are trained on the host by scikit-learn and exported to flat arrays by
// lsml_b200/core/regressors.py.  Arithmetic follows scikit-learn so that velocities match:
//   StandardScaler.transform   (x - mean_) / scale_                       in float64
//   LinearRegression.predict   sum_f x_f * coef_f (f ascending) + intercept in float64
//   RandomForestRegressor      X cast to float32; per tree descend `x[feature] <= threshold`
//                              (threshold pre-rounded DOWN to float32 on the host, which is
//                              exactly equivalent to sklearn's float32-vs-float64 compare);
//                              leaf values summed in tree order in float64, then / n_trees.
//   MLPRegressor               relu/tanh/logistic/identity hidden layers, identity output.
//
// Where the features come from (FeatSrc): either a materialised (B, F) matrix X (the C ABI's
// lsml_*_predict, FeatureMap / fit), or -- the evolution loop -- the feature PROGRAM itself
// (band_features.cuh): every kernel then gathers its voxel's columns from the image planes and
// the per-item scalars, standardises them in registers and never reads or writes X
// ("algorithmic minimum: not materialised, fused into K4").  Both sources yield
// bit-identical values: X[b][f] IS band_feature(b, f).
//
// Forest layout: 32-byte "super nodes" holding an internal node AND its two children (two tree
// levels per fetch; one 256-bit load per lane and level pair):
//   { float thr[3] (node, left child, right child);
//     u32 features (3 x 8 bit) | flags << 24: bit 0/1 = left/right child is a leaf,
//                                             bits 2..5 = grandchild LL, LR, RL, RR is a leaf;
//     u32 ref[4]: per grandchild the index of its super node, or of its leaf value; for a leaf
//                 child, ref[2 * side] is that child's leaf index }
// plus one float64 array of leaf values.  The traversal is latency / L1-tag bound pointer chasing
// (not HBM-bound); one thread per band voxel with the voxel's float32 features staged in shared
// memory as [feature][thread] (conflict-free).  Super nodes halve the number of dependent,
// fully divergent fetches per tree walk compared with the 8-byte single nodes of round 1.
// Measured (round 2, tools/forest_probe.py: 285 k band voxels x 16 features, RF-100 of 599-node
// trees, B200): super nodes 531 us; 8-byte single nodes, one level per fetch, 582 us; every lane
// advancing through the forest at its own pace (no idling on the warp's deepest lane) 568 us;
// the current tree staged in shared memory (cp.async, double-buffered) with 2 / 4 voxels per
// thread 560 / 599 us (kept below: forest_staged_kernel, LSML_FOREST_VARIANT=staged2|staged4).
// Two threads per voxel (half of the forest each, leaf references handed over through shared
// memory so that the sum keeps sklearn's order) 704 us.  Six different walks, one time: what
// they share is the number of L1 data-pipe WAVEFRONTS -- a divergent fetch costs one wavefront
// per distinct sector (global) or per bank conflict (shared), ~86 per warp and tree either way.
// ncu (profiles/r2_ncu_forest.txt): l1tex__data_pipe_lsu_wavefronts 54 % of peak over the whole
// launch with 48 % of the warp slots occupied (one wave, ending with its slowest warps), i.e.
// saturated while the SMs are full; issue slots 41 %, 22 of 32 lanes active.  The next step
// would be 4-byte nodes (thresholds replaced by their rank among the forest's thresholds of that
// feature, features by their rank) walked from shared memory: ~half the wavefronts.
#include <cstdlib>
#include "lsml_internal.h"
#include "band_features.cuh"

namespace lsml {

struct Scaler {
    const double* mean;   // may be null => no scaler
    const double* scale;
};

struct FeatSrc {
    const double* X;      // (B, F) row-major, or null => evaluate the program
    int F;
    BandFeatures bf;
};

__device__ __forceinline__ double scale_value(double x, int f, const Scaler& sc) {
    if (sc.mean != nullptr) x = (x - __ldg(sc.mean + f)) / __ldg(sc.scale + f);
    return x;
}

// standardised feature f of band row b (v: the decoded voxel when the program is the source)
__device__ __forceinline__ double scaled(const FeatSrc& src, i64 b, const VoxelRef& v, int f,
                                         const Scaler& sc) {
    const double x = src.X != nullptr ? __ldg(src.X + b * src.F + f) : band_feature(src.bf, v, f);
    return scale_value(x, f, sc);
}

__device__ __forceinline__ VoxelRef row_ref(const FeatSrc& src, i64 b) {
    if (src.X != nullptr) return VoxelRef{0, 0, 0, 0, 0};
    return voxel_ref(src.bf, __ldg(src.bf.band_idx + b));
}

// --------------------------------------------------------------------------------- linear
__global__ void __launch_bounds__(256)
linear_predict_kernel(FeatSrc src, const i64* __restrict__ n_band, Scaler sc,
                      const double* __restrict__ coef, double intercept,
                      double* __restrict__ out) {
    const i64 nb = *n_band;
    for (i64 b = (i64)blockIdx.x * blockDim.x + threadIdx.x; b < nb;
         b += (i64)gridDim.x * blockDim.x) {
        const VoxelRef v = row_ref(src, b);
        double acc = 0.0;
        for (int f = 0; f < src.F; ++f) acc += scaled(src, b, v, f, sc) * __ldg(coef + f);
        out[b] = acc + intercept;
    }
}

// --------------------------------------------------------------------------------- forest
struct __align__(32) SuperNode {
    float thr[3];
    unsigned ff;          // feature0 | featureL << 8 | featureR << 16 | flags << 24
    unsigned ref[4];
};

struct ForestView {
    const SuperNode* nodes;    // super nodes of all trees
    const double* leaves;      // leaf values of all trees
    const unsigned* root_ref;  // [n_trees] root super node, or the leaf index of a single-leaf tree
    const u8* root_is_leaf;    // [n_trees]
    int n_trees;
    double divisor;            // n_estimators for forests, 1 for a single tree
};

__device__ __forceinline__ SuperNode load_node(const SuperNode* p) {
    SuperNode r;
    unsigned long long a, b, c, d;
    asm("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    r.thr[0] = __uint_as_float((unsigned)a); r.thr[1] = __uint_as_float((unsigned)(a >> 32));
    r.thr[2] = __uint_as_float((unsigned)b); r.ff = (unsigned)(b >> 32);
    r.ref[0] = (unsigned)c; r.ref[1] = (unsigned)(c >> 32);
    r.ref[2] = (unsigned)d; r.ref[3] = (unsigned)(d >> 32);
    return r;
}

constexpr int FOREST_THREADS = 128;

// the canonical 8-byte single-node export (regressors.py: pack_trees)
struct StagedForestView {
    const uint2* nodes;        // all trees back to back
    const int* tree_base;      // [n_trees + 1]
    const u8* root_is_leaf;    // [n_trees]
    int n_trees;
    int max_nodes;             // largest tree
    double divisor;
};

__global__ void __launch_bounds__(FOREST_THREADS)
forest_predict_kernel(FeatSrc src, const i64* __restrict__ n_band, Scaler sc, ForestView fv,
                      double* __restrict__ out) {
    extern __shared__ float feat[];   // [F][FOREST_THREADS]
    const i64 nb = *n_band;
    const i64 nblocks = (nb + FOREST_THREADS - 1) / FOREST_THREADS;
    const int F = src.F;
    for (i64 blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const i64 b = blk * FOREST_THREADS + threadIdx.x;
        const bool live = b < nb;
        __syncthreads();
        if (live) {
            const VoxelRef v = row_ref(src, b);
            for (int f = 0; f < F; ++f)
                feat[f * FOREST_THREADS + threadIdx.x] = __double2float_rn(scaled(src, b, v, f, sc));
        }
        __syncthreads();
        if (!live) continue;
        const float* my = feat + threadIdx.x;
        double acc = 0.0;
        for (int t = 0; t < fv.n_trees; ++t) {
            unsigned ref = __ldg(fv.root_ref + t);
            bool leaf = __ldg(fv.root_is_leaf + t) != 0;
            while (!leaf) {
                const SuperNode nd = load_node(fv.nodes + ref);
                const unsigned flags = nd.ff >> 24;
                const bool r0 = !(my[(nd.ff & 0xffu) * FOREST_THREADS] <= nd.thr[0]);
                if ((flags >> (r0 ? 1 : 0)) & 1u) {          // that child is a leaf
                    ref = r0 ? nd.ref[2] : nd.ref[0];
                    break;
                }
                const unsigned f1 = r0 ? (nd.ff >> 16) & 0xffu : (nd.ff >> 8) & 0xffu;
                const bool r1 = !(my[f1 * FOREST_THREADS] <= (r0 ? nd.thr[2] : nd.thr[1]));
                const unsigned lo = r1 ? nd.ref[1] : nd.ref[0], hi = r1 ? nd.ref[3] : nd.ref[2];
                ref = r0 ? hi : lo;
                leaf = (flags >> (2 + (r0 ? 2 : 0) + (r1 ? 1 : 0))) & 1u;
            }
            acc += __ldg(fv.leaves + ref);
        }
        out[b] = acc / fv.divisor;
    }
}

// Forests whose trees are small (RandomForest with max_samples, the configuration of the
// reference's examples: <= ~600 nodes, 4.8 KB per tree): the CTA walks the forest tree by tree
// with the CURRENT tree staged in shared memory (cp.async, double-buffered: tree t+1 streams in
// while tree t is walked) and V voxels per thread, i.e. V independent pointer chases per thread
// to hide the shared-memory latency.  A divergent node fetch then costs the bank conflicts of a
// 64-bit shared load (~3-4 wavefronts per warp) instead of up to 32 L1 tag lookups.  Nodes are
// the canonical 8-byte single nodes (children adjacent; the parent's word says which child is a
// leaf; a leaf holds its float64 value), see lsml_b200/core/regressors.py: pack_trees.
typedef StagedForestView StagedForest;

constexpr int STAGED_THREADS = 256;

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

template <int V>
__global__ void __launch_bounds__(STAGED_THREADS)
forest_staged_kernel(FeatSrc src, const i64* __restrict__ n_band, Scaler sc, StagedForest fv,
                     double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int PER_CTA = V * STAGED_THREADS;
    uint2* tree_buf[2];
    tree_buf[0] = reinterpret_cast<uint2*>(smem_raw);
    tree_buf[1] = tree_buf[0] + fv.max_nodes;
    float* feat = reinterpret_cast<float*>(tree_buf[1] + fv.max_nodes);     // [F][PER_CTA]
    const i64 nb = *n_band;
    const i64 nblocks = (nb + PER_CTA - 1) / PER_CTA;
    const int F = src.F;
    for (i64 blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        __syncthreads();                          // previous block's walks are done with smem
        // tree 0 starts streaming while the features are gathered
        {
            const int b0 = __ldg(fv.tree_base), n0 = __ldg(fv.tree_base + 1) - b0;
            for (int k = threadIdx.x; k < n0; k += STAGED_THREADS) cp_async8(tree_buf[0] + k, fv.nodes + b0 + k);
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const i64 b = blk * PER_CTA + v * STAGED_THREADS + threadIdx.x;
            if (b < nb) {
                const VoxelRef vr = row_ref(src, b);
                for (int f = 0; f < F; ++f)
                    feat[f * PER_CTA + v * STAGED_THREADS + threadIdx.x] =
                        __double2float_rn(scaled(src, b, vr, f, sc));
            }
        }
        double acc[V];
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] = 0.0;
        cp_async_commit_wait_all();
        __syncthreads();
        for (int t = 0; t < fv.n_trees; ++t) {
            const uint2* tree = tree_buf[t & 1];
            if (t + 1 < fv.n_trees) {             // stream the next tree into the other buffer
                const int b1 = __ldg(fv.tree_base + t + 1), n1 = __ldg(fv.tree_base + t + 2) - b1;
                uint2* dst = tree_buf[(t + 1) & 1];
                for (int k = threadIdx.x; k < n1; k += STAGED_THREADS) cp_async8(dst + k, fv.nodes + b1 + k);
            }
            const bool root_leaf = __ldg(fv.root_is_leaf + t) != 0;
            int idx[V];
            unsigned live = 0u;                   // chains still descending
#pragma unroll
            for (int v = 0; v < V; ++v) {
                idx[v] = 0;
                const i64 b = blk * PER_CTA + v * STAGED_THREADS + threadIdx.x;
                if (b < nb && !root_leaf) live |= 1u << v;
            }
            while (live) {
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    if (!((live >> v) & 1u)) continue;
                    const uint2 nd = tree[idx[v]];
                    const unsigned meta = nd.y;
                    const float x = feat[(meta >> 24) * PER_CTA + v * STAGED_THREADS + threadIdx.x];
                    const bool right = !(x <= __uint_as_float(nd.x));
                    idx[v] = (int)(meta & 0x3fffffu) + (right ? 1 : 0);
                    if (right ? ((meta >> 22) & 1u) : ((meta >> 23) & 1u)) live &= ~(1u << v);
                }
            }
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const uint2 lv = tree[idx[v]];    // (a dead lane reads node 0: harmless)
                acc[v] += __hiloint2double((int)lv.y, (int)lv.x);
            }
            cp_async_commit_wait_all();
            __syncthreads();                      // tree t+1 complete, tree t's buffer free
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const i64 b = blk * PER_CTA + v * STAGED_THREADS + threadIdx.x;
            if (b < nb) out[b] = acc[v] / fv.divisor;
        }
    }
}

// --------------------------------------------------------------------------- ranked forest
// Rank-quantised forest walked from shared memory (lsml_b200/core/regressors.py: pack_ranked).
// A tree only asks `x[f] <= thr`; with S_f the sorted distinct float32 thresholds of feature f
// and rank_f(x) = #{s in S_f : s < x}, that is `rank_f(x) <= j` for the node holding S_f[j].  A
// voxel's features are reduced ONCE to ranks (one binary search per feature), an internal node
// is ONE 32-bit word, fields from the top
//     feature (fbits) | rank j (rbits) | offset of the right child in words | bit 1 / bit 0: the
//     right / left child is a leaf
// trees are stored in preorder (left child = next word, a leaf = its float64 value in two
// words), and the whole forest is ~3.7 KB per 300-leaf tree: a CTA stages GROUPS of ~20 trees in
// shared memory and walks them for its 512 (x V) voxels, the float64 sum keeping sklearn's tree
// order.  A node visit is two shared-memory loads -- the node (divergent: ~3-4 bank-conflict
// wavefronts per warp instead of up to 32 L1 tag lookups of a 32-byte global node) and the
// voxel's rank of that feature ([feature][voxel] layout: conflict-free).  The voxel's rank is
// stored as `feature << (32 - fbits) | rank << (32 - fbits - rbits)`: "go right" is then ONE
// unsigned comparison of that word with the node word (same top field, the fields below the rank
// cannot flip the result), and `node & offset mask` is the byte offset of the right child.
// ncu of the first version (profiles/r2b_ncu_forest_ranked.txt): neither shared-memory
// wavefronts (42 %) nor issue slots (64 %) bound the walk but the INTEGER ALU pipe (half rate:
// LOP3 / SHF / SEL / ISETP cost two cycles per warp; 8 of the 13 instructions per level), which is
// why doubling the resident warps changed nothing.  The walk below spends 4-5 ALU-pipe and 3-4
// FMA-pipe (IMAD / IMAD.HI) instructions per level instead.
struct RankedView {
    const unsigned* words;
    const int* tree_start;     // [n_trees + 1], word offsets (multiples of 4)
    const u8* root_is_leaf;    // [n_trees]
    const float* thr_tab;      // S_0 | S_1 | ...
    const int* thr_off;        // [F + 1]
    int n_trees, fbits, rbits;
    double divisor;
};

constexpr int RK_THREADS = 512;
constexpr int RK_MAX_GROUP = 128;          // trees per staged group
constexpr int RK_SMEM_PER_CTA = 112 * 1024;

// #{s in S_f : s < x}, NaN -> |S_f| (ranks above every threshold: `!(x <= thr)` goes right)
__device__ __forceinline__ unsigned rank_of(const float* __restrict__ tab, int lo, int hi, float x) {
    if (x != x) return (unsigned)(hi - lo);
    const int first = lo;
    int len = hi - lo;
    while (len > 0) {
        const int half = len >> 1;
        if (__ldg(tab + lo + half) < x) { lo += half + 1; len -= half + 1; }
        else len = half;
    }
    return (unsigned)(lo - first);
}

__device__ __forceinline__ unsigned lds32(unsigned saddr) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

__device__ __forceinline__ unsigned mad32(unsigned a, unsigned b, unsigned c) {
    unsigned d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// ---- bulk (TMA) copy of a tree group: one thread issues it, everybody waits on the mbarrier ----
__device__ __forceinline__ void rk_mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void rk_mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rk_mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "RK_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra RK_DONE;\n"
        "bra RK_WAIT;\n"
        "RK_DONE:\n"
        "}\n" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void rk_bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int V, int CTAS, int THREADS>
__global__ void __launch_bounds__(THREADS, CTAS)
forest_ranked_kernel(FeatSrc src, const i64* __restrict__ n_band, Scaler sc, RankedView rv,
                     double* __restrict__ out, int cap_words) {
    extern __shared__ __align__(16) unsigned rk_smem[];
    __shared__ int s_base[RK_MAX_GROUP];
    __shared__ u8 s_rootleaf[RK_MAX_GROUP];
    __shared__ int s_t1;
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ unsigned char s_itemf[64];            // the per-ITEM columns of the feature program
    __shared__ int s_nitemf;
    constexpr int N = V * THREADS;
    const int F = src.F;
    unsigned* const ranks = rk_smem;                 // [F][N], feature | rank (node word fields)
    unsigned* const trees = rk_smem + (size_t)F * N; // [cap_words]
    const int fshift = 32 - rv.fbits;                // feature field: top fbits bits
    const int rshift = fshift - rv.rbits;            // rank field below it
    const unsigned fmul = 1u << rv.fbits;            // umulhi(word, fmul) = feature
    const unsigned omask4 = (1u << rshift) - 4u;     // word offset field << 2 = byte offset
    // 32-bit shared addresses: this thread's column of the rank table, the tree buffer
    const unsigned my_sa = (unsigned)__cvta_generic_to_shared(ranks + threadIdx.x);
    const unsigned trees_sa = (unsigned)__cvta_generic_to_shared(trees);
    const unsigned bar_sa = (unsigned)__cvta_generic_to_shared(&s_bar);
    unsigned parity = 0u;
    if (threadIdx.x == 0) {
        rk_mbar_init(bar_sa, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // columns that have ONE value per item (band mean / std, size, moments, ...): a warp whose
        // 32 voxels belong to one item ranks each of them once, on one lane
        int cnt = 0;
        if (src.X == nullptr)
            for (int f = 0; f < src.F; ++f) {
                const int kind = src.bf.prog.kind[f];
                if (kind != FK_PLANE && kind != FK_DIST_COM && kind != FK_COM_RAY) s_itemf[cnt++] = (unsigned char)f;
            }
        s_nitemf = cnt;
    }
    __syncthreads();
    const int n_itemf = s_nitemf;
    const i64 nb = *n_band;
    const i64 nchunks = (nb + N - 1) / N;
    for (i64 chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        // ---- features -> ranks: one binary search per feature in the threshold tables (global
        // memory, L2-resident; staging them through the tree buffer first was measured and is
        // no faster: the searches of one CTA overlap with the other CTA's walks)
        __syncthreads();                             // the previous chunk's walks are done
        unsigned valid = 0u;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const int slot = v * THREADS + threadIdx.x;
            const i64 b = chunk * N + slot;
            const bool live = b < nb;
            const VoxelRef vr = row_ref(src, live ? b : 0);
            int same = 0;                            // all 32 voxels of the warp live, in one item
            if (n_itemf > 0) __match_all_sync(0xffffffffu, live ? vr.item : (i64)-1 - (threadIdx.x & 31), &same);
            if (live) {
                valid |= 1u << v;
                for (int f = 0; f < F; ++f) {
                    if (same) {
                        const int kind = src.bf.prog.kind[f];
                        if (kind != FK_PLANE && kind != FK_DIST_COM && kind != FK_COM_RAY) continue;
                    }
                    const float x = __double2float_rn(scaled(src, b, vr, f, sc));
                    const unsigned r = rank_of(rv.thr_tab, __ldg(rv.thr_off + f),
                                               __ldg(rv.thr_off + f + 1), x);
                    ranks[f * N + slot] = ((unsigned)f << fshift) | (r << rshift);
                }
            }
            if (same) {
                // lane i ranks the i-th per-item column (all searches side by side instead of
                // one after the other on every lane), then the words are handed round
                const int lane = threadIdx.x & 31;
                for (int base = 0; base < n_itemf; base += 32) {
                    const int cnt = min(32, n_itemf - base);
                    unsigned word = 0u, fcol = 0u;
                    if (lane < cnt) {
                        fcol = s_itemf[base + lane];
                        const float x = __double2float_rn(scaled(src, b, vr, (int)fcol, sc));
                        const unsigned r = rank_of(rv.thr_tab, __ldg(rv.thr_off + fcol),
                                                   __ldg(rv.thr_off + fcol + 1), x);
                        word = (fcol << fshift) | (r << rshift);
                    }
                    for (int j = 0; j < cnt; ++j) {
                        const unsigned w = __shfl_sync(0xffffffffu, word, j);
                        const unsigned fj = __shfl_sync(0xffffffffu, fcol, j);
                        ranks[fj * N + slot] = w;
                    }
                }
            }
        }
        double acc[V];
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] = 0.0;
        int t0 = 0;
        while (t0 < rv.n_trees) {
            // group [t0, t1): as many whole trees as the buffer holds (warp 0 looks 32 trees
            // ahead per step; tree_start is increasing, so the fitting trees are a prefix)
            const int gbase = __ldg(rv.tree_start + t0);
            if (threadIdx.x < 32) {
                int t1w = t0 + 1;
                for (;;) {
                    const int t = t1w + (int)threadIdx.x;
                    const bool ok = t < rv.n_trees && t - t0 < RK_MAX_GROUP &&
                                    __ldg(rv.tree_start + t + 1) - gbase <= cap_words;
                    const unsigned bal = __ballot_sync(0xffffffffu, ok);
                    if (bal != 0xffffffffu) { t1w += __ffs((int)~bal) - 1; break; }
                    t1w += 32;
                }
                if (threadIdx.x == 0) s_t1 = t1w;
            }
            __syncthreads();                         // the previous group's walks / the ranks are done
            const int t1 = s_t1;
            const int gwords = __ldg(rv.tree_start + t1) - gbase;
            if (threadIdx.x == 0) {
                // the buffer was read through the generic proxy until the barrier above; the bulk
                // copy writes it through the async proxy
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                rk_mbar_expect_tx(bar_sa, (unsigned)gwords * 4u);
                rk_bulk_g2s(trees_sa, rv.words + gbase, (unsigned)gwords * 4u, bar_sa);
            }
            for (int t = t0 + threadIdx.x; t < t1; t += THREADS) {
                s_base[t - t0] = (__ldg(rv.tree_start + t) - gbase) * 4;          // bytes
                s_rootleaf[t - t0] = __ldg(rv.root_is_leaf + t);
            }
            rk_mbar_wait(bar_sa, parity);            // the group has landed
            parity ^= 1u;
            __syncthreads();                         // ... and s_base / s_rootleaf are written
            for (int t = 0; t < t1 - t0; ++t) {
                const unsigned root = trees_sa + (unsigned)s_base[t];
                if (V == 1) {
                    unsigned ia = root;              // shared address of the chain's node
                    bool live = (valid & 1u) && !s_rootleaf[t];
                    while (live) {
                        const unsigned n = lds32(ia);
                        // row `feature` of the rank table (N * 4 bytes per row)
                        const unsigned w = lds32(mad32(__umulhi(n, fmul), N * 4, my_sa));
                        const bool right = w > n;
                        if (right) ia += n & omask4; else ia += 4u;    // left child: next word
                        live = !(n & (right ? 2u : 1u));
                    }
                    // (a dead slot reads two node words as a value that is never stored)
                    acc[0] += __hiloint2double((int)lds32(ia + 4u), (int)lds32(ia));
                } else {
                    // V independent chains per thread, walked in lockstep without branches so
                    // that their shared-memory round trips overlap; a finished chain stays on
                    // its last internal node (re-reading it is harmless) and remembers its leaf
                    unsigned ia[V], leaf[V];
                    unsigned live = s_rootleaf[t] ? 0u : valid;
#pragma unroll
                    for (int v = 0; v < V; ++v) { ia[v] = root; leaf[v] = root; }
                    while (live) {
                        unsigned n[V], w[V];
#pragma unroll
                        for (int v = 0; v < V; ++v) n[v] = lds32(ia[v]);
#pragma unroll
                        for (int v = 0; v < V; ++v)
                            w[v] = lds32(mad32(__umulhi(n[v], fmul), N * 4, my_sa + v * THREADS * 4));
#pragma unroll
                        for (int v = 0; v < V; ++v) {
                            const bool right = w[v] > n[v];
                            const unsigned nx = right ? ia[v] + (n[v] & omask4) : ia[v] + 4u;
                            const bool on = (live >> v) & 1u;
                            const bool ends = (n[v] & (right ? 2u : 1u)) != 0u;
                            if (on && ends) { leaf[v] = nx; live &= ~(1u << v); }
                            if (on && !ends) ia[v] = nx;
                        }
                    }
#pragma unroll
                    for (int v = 0; v < V; ++v)
                        acc[v] += __hiloint2double((int)lds32(leaf[v] + 4u), (int)lds32(leaf[v]));
                }
            }
            t0 = t1;
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const i64 b = chunk * N + v * THREADS + threadIdx.x;
            if (b < nb) out[b] = acc[v] / rv.divisor;
        }
    }
}

// --------------------------------------------------------------------------------- MLP
// One thread per band voxel; weights are read through the read-only path (L1-resident: a
// (F x H) float64 layer with F=16, H=100 is 12.8 KB).  Activations ping-pong in shared memory
// as [unit][thread].  act: 0 identity, 1 relu, 2 tanh, 3 logistic.  float64 on purpose: the
// reference's MLPRegressor.predict is float64 and the velocities must agree to 1e-5 AFTER up to
// hundreds of iterations; at F <= 64 and one hidden layer the dense work is ~2 x 10^4 flops per
// voxel next to ~150 bytes of gathers -- tensor cores have nothing to win here.
struct MlpView {
    int n_layers;              // number of weight matrices
    int sizes[9];              // sizes[0] = F, sizes[n_layers] = 1
    const double* W[8];        // W[l] row-major (sizes[l], sizes[l+1]) as sklearn's coefs_
    const double* b[8];
    int act;
    int max_width;
};

constexpr int MLP_THREADS = 64;

__device__ __forceinline__ double mlp_act(double x, int act) {
    if (act == 1) return x > 0.0 ? x : 0.0;          // np.maximum(x, 0)
    if (act == 2) return tanh(x);
    if (act == 3) return 1.0 / (1.0 + exp(-x));      // scipy.special.expit
    return x;
}

__global__ void __launch_bounds__(MLP_THREADS)
mlp_predict_kernel(FeatSrc src, const i64* __restrict__ n_band, Scaler sc, MlpView mv,
                   double* __restrict__ out) {
    extern __shared__ double actbuf[];   // 2 x [max_width][MLP_THREADS]
    double* cur = actbuf;
    double* nxt = actbuf + (size_t)mv.max_width * MLP_THREADS;
    const i64 nb = *n_band;
    const int F = mv.sizes[0];
    for (i64 b = (i64)blockIdx.x * MLP_THREADS + threadIdx.x; b < nb;
         b += (i64)gridDim.x * MLP_THREADS) {
        const VoxelRef v = row_ref(src, b);
        for (int f = 0; f < F; ++f) cur[f * MLP_THREADS + threadIdx.x] = scaled(src, b, v, f, sc);
        double* a = cur; double* o = nxt;
        for (int l = 0; l < mv.n_layers; ++l) {
            const int nin = mv.sizes[l], nout = mv.sizes[l + 1];
            for (int j = 0; j < nout; ++j) {
                double acc = 0.0;
                for (int i = 0; i < nin; ++i)
                    acc += a[i * MLP_THREADS + threadIdx.x] * __ldg(mv.W[l] + (i64)i * nout + j);
                acc = acc + __ldg(mv.b[l] + j);
                o[j * MLP_THREADS + threadIdx.x] = (l + 1 < mv.n_layers) ? mlp_act(acc, mv.act) : acc;
            }
            double* t = a; a = o; o = t;
        }
        out[b] = a[threadIdx.x];
    }
}

// ------------------------------------------------------------------------- MLP on tensor cores
// The dense layers as FP64 tensor-core MMAs (`mma.sync.m8n8k4.f64`: DMMA in the SASS).  float64
// keeps the reference's precision (MLPRegressor.predict is float64; the level sets must agree
// to 1e-5 after hundreds of iterations), the accumulation is fused multiply-add in k-blocks of
// four instead of sklearn's BLAS order: same tolerance as the scalar kernel (1e-11 relative).
// A warp owns a tile of 8 band voxels: its activations live in shared memory as [8][stride]
// (stride = width rounded up to 16, + 4: the A-fragment reads of a half-warp then hit 16
// distinct 8-byte banks), ping-pong between layers; per layer and block of 8 outputs the warp
// runs ceil(nin / 4) MMAs, adds the bias, applies the activation and stores the D fragment as the
// next layer's input (columns beyond the layer's width are stored as zeros, so the next k-loop can
// run to a multiple of four).  Weights come through the read-only path as B fragments (a
// (16 x 100) layer is 12.8 KB: L1-resident).  Fragment layouts (PTX ISA, m8n8k4 .f64): A[row =
// lane / 4][k = lane % 4], B[k = lane % 4][col = lane / 4], C / D[row = lane / 4][col = 2 * (lane
// % 4) + {0, 1}].
constexpr int DM_WARPS = 4;
constexpr int DM_THREADS = DM_WARPS * 32;

__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(DM_THREADS)
mlp_dmma_kernel(FeatSrc src, const i64* __restrict__ n_band, Scaler sc, MlpView mv, int stride,
                double* __restrict__ out) {
    extern __shared__ double dm_act[];   // [warp][2][8][stride]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = lane >> 2, col = lane & 3;
    double* cur = dm_act + (size_t)warp * 2 * 8 * stride;
    double* nxt = cur + 8 * stride;
    const i64 nb = *n_band;
    const i64 ntiles = (nb + 7) / 8;
    const int F = mv.sizes[0];
    for (i64 tile = (i64)blockIdx.x * DM_WARPS + warp; tile < ntiles; tile += (i64)gridDim.x * DM_WARPS) {
        const i64 b = tile * 8 + row;
        const bool live = b < nb;
        // inputs: the four lanes of a voxel's row gather its features k = col, col + 4, ...;
        // the row is zero-padded to a multiple of four
        {
            VoxelRef v = row_ref(src, live ? b : 0);
            const int fpad = (F + 3) & ~3;
            for (int f = col; f < fpad; f += 4)
                cur[row * stride + f] = (live && f < F) ? scaled(src, b, v, f, sc) : 0.0;
        }
        __syncwarp();
        double* a_in = cur; double* a_out = nxt;
        for (int l = 0; l < mv.n_layers; ++l) {
            const int nin = mv.sizes[l], nout = mv.sizes[l + 1];
            const double* __restrict__ W = mv.W[l];
            const double* __restrict__ B = mv.b[l];
            const bool last = l + 1 == mv.n_layers;
            for (int n0 = 0; n0 < nout; n0 += 8) {
                double d0 = 0.0, d1 = 0.0;
                const int bn = n0 + row;                       // this lane's B column
                for (int k0 = 0; k0 < nin; k0 += 4) {
                    const double a = a_in[row * stride + k0 + col];
                    const int bk = k0 + col;
                    const double w = (bk < nin && bn < nout) ? __ldg(W + (i64)bk * nout + bn) : 0.0;
                    dmma_8x8x4(d0, d1, a, w);
                }
                const int c0 = n0 + 2 * col, c1 = c0 + 1;
                double y0 = 0.0, y1 = 0.0;
                if (c0 < nout) { y0 = d0 + __ldg(B + c0); if (!last) y0 = mlp_act(y0, mv.act); }
                if (c1 < nout) { y1 = d1 + __ldg(B + c1); if (!last) y1 = mlp_act(y1, mv.act); }
                a_out[row * stride + c0] = y0;
                a_out[row * stride + c1] = y1;
            }
            __syncwarp();
            double* t = a_in; a_in = a_out; a_out = t;
        }
        if (live && col == 0) out[b] = a_in[row * stride];
        __syncwarp();
    }
}

// --------------------------------------------------------------------------------- launchers
static FeatSrc matrix_source(const double* X, int F) {
    FeatSrc src;
    src.X = X; src.F = F;
    src.bf = BandFeatures{};
    return src;
}
static FeatSrc program_source(const BandFeatures& bf) {
    FeatSrc src;
    src.X = nullptr; src.F = bf.prog.nfeat; src.bf = bf;
    return src;
}

static int linear_launch(const FeatSrc& src, const i64* n_band, const double* mean,
                         const double* scale, const double* coef, double intercept, double* out,
                         cudaStream_t s) {
    Scaler sc{mean, scale};
    linear_predict_kernel<<<LSML_SMS * 8, 256, 0, s>>>(src, n_band, sc, coef, intercept, out);
    LSML_LAUNCH_CHECK("linear_predict_kernel");
    return 0;
}

// variant selection (probes): LSML_FOREST_VARIANT = "staged2" / "staged4" runs the shared-memory
// kernel with 2 / 4 voxels per thread instead of the super-node kernel
static int forest_variant() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("LSML_FOREST_VARIANT");
        v = 0;
        if (e && e[0] == 'g') v = 1;
        if (e && e[0] == 's') v = (e[6] == '4') ? 4 : 2;
    }
    return v;
}

template <int V>
static int staged_launch(const FeatSrc& src, const i64* n_band, const Scaler& sc,
                         const StagedForest& fv, double* out, cudaStream_t s, bool* done) {
    const size_t smem = 2 * (size_t)fv.max_nodes * sizeof(uint2) +
                        (size_t)src.F * V * STAGED_THREADS * sizeof(float);
    *done = false;
    if (smem > 100 * 1024) return 0;              // would leave < 2 CTAs per SM: not worth it
    static size_t configured[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (smem > 48 * 1024 && smem > configured[V]) {
        LSML_CUDA(cudaFuncSetAttribute(forest_staged_kernel<V>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[V] = smem;
    }
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    if (per_sm > 2048 / STAGED_THREADS) per_sm = 2048 / STAGED_THREADS;
    forest_staged_kernel<V><<<LSML_SMS * per_sm, STAGED_THREADS, smem, s>>>(src, n_band, sc, fv, out);
    LSML_LAUNCH_CHECK("forest_staged_kernel");
    *done = true;
    return 0;
}

static int forest_launch(const FeatSrc& src, const i64* n_band, const double* mean,
                         const double* scale, const void* nodes, const double* leaves,
                         const unsigned* root_ref, const u8* root_is_leaf, int n_trees,
                         double divisor, const void* nodes8, const int* tree_base, int max_nodes,
                         double* out, cudaStream_t s) {
    if (src.F > 255) { set_error("forest: at most 255 features"); return 1; }
    const int variant = forest_variant();
    if (nodes8 != nullptr && tree_base != nullptr && max_nodes > 0 && max_nodes <= 2048 &&
        (variant == 2 || variant == 4)) {
        const Scaler sc8{mean, scale};
        const StagedForest sf{static_cast<const uint2*>(nodes8), tree_base, root_is_leaf, n_trees,
                              max_nodes, divisor};
        bool done = false;
        const int st = variant == 4 ? staged_launch<4>(src, n_band, sc8, sf, out, s, &done)
                                    : staged_launch<2>(src, n_band, sc8, sf, out, s, &done);
        if (st || done) return st;
    }
    if ((reinterpret_cast<uintptr_t>(nodes) & 31u) != 0) {
        set_error("forest: super nodes must be 32-byte aligned");
        return 1;
    }
    Scaler sc{mean, scale};
    ForestView fv{static_cast<const SuperNode*>(nodes), leaves, root_ref, root_is_leaf, n_trees,
                  divisor};
    const size_t smem = (size_t)src.F * FOREST_THREADS * sizeof(float);
    if (smem > 48 * 1024)
        LSML_CUDA(cudaFuncSetAttribute(forest_predict_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    forest_predict_kernel<<<LSML_SMS * 16, FOREST_THREADS, smem, s>>>(src, n_band, sc, fv, out);
    LSML_LAUNCH_CHECK("forest_predict_kernel");
    return 0;
}

template <int V, int CTAS, int THREADS = RK_THREADS>
static int ranked_launch_v(const FeatSrc& src, const i64* n_band, const Scaler& sc,
                           const RankedView& rv, double* out, size_t smem, cudaStream_t s) {
    static size_t configured = 0;
    if (smem > configured) {
        LSML_CUDA(cudaFuncSetAttribute(forest_ranked_kernel<V, CTAS, THREADS>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const size_t rank_bytes = (size_t)src.F * V * THREADS * sizeof(unsigned);
    const int cap_words = (int)((smem - rank_bytes) / sizeof(unsigned)) & ~3;
    forest_ranked_kernel<V, CTAS, THREADS><<<LSML_SMS * CTAS, THREADS, smem, s>>>(src, n_band, sc, rv, out,
                                                                           cap_words);
    LSML_LAUNCH_CHECK("forest_ranked_kernel");
    return 0;
}

// probes: LSML_FOREST_RK = "<voxels per thread><CTAs per SM>", e.g. 12 (the default), 22, 13
static int ranked_override() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("LSML_FOREST_RK");
        v = (e && e[0] >= '1' && e[0] <= '2' && e[1] >= '1' && e[1] <= '3') ? (e[0] - '0') * 10 + (e[1] - '0') : 0;
        if (e && e[0] == 't') v = 100 + (e[1] - '0');        // t2 / t1: 1024 threads, 2 / 1 CTAs per SM
    }
    return v;
}

static int ranked_launch(const FeatSrc& src, const i64* n_band, const double* mean,
                         const double* scale, const RankedArgs& a, double* out, cudaStream_t s) {
    if (a.nfeat != src.F) { set_error("ranked forest: exported for %d features, source has %d", a.nfeat, src.F); return 1; }
    if (a.n_trees < 1 || a.fbits < 1 || a.rbits < 1 || a.fbits + a.rbits > 28 ||
        src.F > (1 << a.fbits) || a.max_tree_words < 2) {
        set_error("ranked forest: bad layout parameters");
        return 1;
    }
    if ((reinterpret_cast<uintptr_t>(a.words) & 15u) != 0) {
        set_error("ranked forest: words must be 16-byte aligned");
        return 1;
    }
    const Scaler sc{mean, scale};
    const RankedView rv{a.words, a.tree_start, a.root_is_leaf, a.thr_tab, a.thr_off,
                        a.n_trees, a.fbits, a.rbits, a.divisor};
    const size_t tree_bytes = (size_t)a.max_tree_words * sizeof(unsigned);
    const size_t want_trees = tree_bytes > 32 * 1024 ? tree_bytes : 32 * 1024;
    // shared memory per CTA with 3 / 2 / 1 CTAs per SM (228 KB per SM, 1 KB reserved per CTA)
    const size_t three = 74 * 1024, two = RK_SMEM_PER_CTA - 1024, one = 224 * 1024;
    const size_t rank1 = (size_t)src.F * RK_THREADS * sizeof(unsigned);
    const int ov = ranked_override();
    if (ov == 102 && 2 * rank1 + want_trees <= two) return ranked_launch_v<1, 2, 1024>(src, n_band, sc, rv, out, two, s);
    if (ov == 101 && 2 * rank1 + tree_bytes <= one) return ranked_launch_v<1, 1, 1024>(src, n_band, sc, rv, out, one, s);
    if (ov == 13 && rank1 + want_trees <= three) return ranked_launch_v<1, 3>(src, n_band, sc, rv, out, three, s);
    if (ov == 23 && 2 * rank1 + want_trees <= three) return ranked_launch_v<2, 3>(src, n_band, sc, rv, out, three, s);
    if (ov == 22 && 2 * rank1 + want_trees <= two) return ranked_launch_v<2, 2>(src, n_band, sc, rv, out, two, s);
    if (ov == 21 && 2 * rank1 + tree_bytes <= one) return ranked_launch_v<2, 1>(src, n_band, sc, rv, out, one, s);
    // default: one CTA of 1024 threads per SM with the whole 224 KB -- the fewest, largest tree
    // groups (43 trees of the bench's forest at F = 16: 3 groups per chunk instead of 5 with two
    // 512-thread CTAs; measured 329 vs 344 us on the probe state)
    const size_t big_trees = tree_bytes > 64 * 1024 ? tree_bytes : 64 * 1024;
    if (ov == 0 && 2 * rank1 + big_trees <= one)
        return ranked_launch_v<1, 1, 1024>(src, n_band, sc, rv, out, one, s);
    if (rank1 + want_trees <= two) return ranked_launch_v<1, 2>(src, n_band, sc, rv, out, two, s);
    if (rank1 + tree_bytes <= one) return ranked_launch_v<1, 1>(src, n_band, sc, rv, out, one, s);
    set_error("ranked forest: %d features and trees of %d words do not fit shared memory",
              src.F, a.max_tree_words);
    return 1;
}

static int mlp_launch(const FeatSrc& src, const i64* n_band, const double* mean,
                      const double* scale, int n_layers, const int* sizes, const double* const* W,
                      const double* const* b, int act, double* out, cudaStream_t s) {
    if (n_layers < 1 || n_layers > 8) { set_error("mlp: 1..8 layers supported"); return 1; }
    if (sizes[0] != src.F) { set_error("mlp: %d inputs, feature source has %d", sizes[0], src.F); return 1; }
    Scaler sc{mean, scale};
    MlpView mv;
    mv.n_layers = n_layers; mv.act = act; mv.max_width = 0;
    for (int l = 0; l <= n_layers; ++l) {
        mv.sizes[l] = sizes[l];
        if (sizes[l] > mv.max_width) mv.max_width = sizes[l];
    }
    for (int l = 0; l < n_layers; ++l) { mv.W[l] = W[l]; mv.b[l] = b[l]; }
    // tensor-core path (LSML_MLP_VARIANT=scalar selects the one-thread-per-voxel kernel below)
    static const bool scalar_only = [] { const char* e = getenv("LSML_MLP_VARIANT"); return e && e[0] == 's'; }();
    const int stride = ((mv.max_width + 15) & ~15) + 4;
    const size_t dsmem = (size_t)DM_WARPS * 2 * 8 * stride * sizeof(double);
    if (!scalar_only && dsmem <= 200 * 1024) {
        static size_t configured = 0;
        if (dsmem > 48 * 1024 && dsmem > configured) {
            LSML_CUDA(cudaFuncSetAttribute(mlp_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)dsmem));
            configured = dsmem;
        }
        mlp_dmma_kernel<<<LSML_SMS * 8, DM_THREADS, dsmem, s>>>(src, n_band, sc, mv, stride, out);
        LSML_LAUNCH_CHECK("mlp_dmma_kernel");
        return 0;
    }
    const size_t smem = 2 * (size_t)mv.max_width * MLP_THREADS * sizeof(double);
    if (smem > 200 * 1024) { set_error("mlp: layer width %d too large", mv.max_width); return 1; }
    if (smem > 48 * 1024)
        LSML_CUDA(cudaFuncSetAttribute(mlp_predict_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mlp_predict_kernel<<<LSML_SMS * 4, MLP_THREADS, smem, s>>>(src, n_band, sc, mv, out);
    LSML_LAUNCH_CHECK("mlp_predict_kernel");
    return 0;
}

int linear_predict(const double* X, const i64* n_band, int F, const double* mean,
                   const double* scale, const double* coef, double intercept, double* out,
                   cudaStream_t s) {
    return linear_launch(matrix_source(X, F), n_band, mean, scale, coef, intercept, out, s);
}
int linear_predict_band(const BandFeatures& bf, const i64* n_band, const double* mean,
                        const double* scale, const double* coef, double intercept, double* out,
                        cudaStream_t s) {
    return linear_launch(program_source(bf), n_band, mean, scale, coef, intercept, out, s);
}

int forest_predict(const double* X, const i64* n_band, int F, const double* mean,
                   const double* scale, const void* nodes, const double* leaves,
                   const unsigned* root_ref, const u8* root_is_leaf, int n_trees, double divisor,
                   const void* nodes8, const int* tree_base, int max_nodes, double* out,
                   cudaStream_t s) {
    return forest_launch(matrix_source(X, F), n_band, mean, scale, nodes, leaves, root_ref,
                         root_is_leaf, n_trees, divisor, nodes8, tree_base, max_nodes, out, s);
}
int forest_predict_band(const BandFeatures& bf, const i64* n_band, const double* mean,
                        const double* scale, const void* nodes, const double* leaves,
                        const unsigned* root_ref, const u8* root_is_leaf, int n_trees,
                        double divisor, const void* nodes8, const int* tree_base, int max_nodes,
                        double* out, cudaStream_t s) {
    return forest_launch(program_source(bf), n_band, mean, scale, nodes, leaves, root_ref,
                         root_is_leaf, n_trees, divisor, nodes8, tree_base, max_nodes, out, s);
}

int forest_ranked_predict(const double* X, const i64* n_band, int F, const double* mean,
                          const double* scale, const RankedArgs& a, double* out, cudaStream_t s) {
    return ranked_launch(matrix_source(X, F), n_band, mean, scale, a, out, s);
}
int forest_ranked_predict_band(const BandFeatures& bf, const i64* n_band, const double* mean,
                               const double* scale, const RankedArgs& a, double* out,
                               cudaStream_t s) {
    return ranked_launch(program_source(bf), n_band, mean, scale, a, out, s);
}

int mlp_predict(const double* X, const i64* n_band, const double* mean, const double* scale,
                int n_layers, const int* sizes, const double* const* W, const double* const* b,
                int act, double* out, cudaStream_t s) {
    return mlp_launch(matrix_source(X, sizes[0]), n_band, mean, scale, n_layers, sizes, W, b, act,
                      out, s);
}
int mlp_predict_band(const BandFeatures& bf, const i64* n_band, const double* mean,
                     const double* scale, int n_layers, const int* sizes, const double* const* W,
                     const double* const* b, int act, double* out, cudaStream_t s) {
    return mlp_launch(program_source(bf), n_band, mean, scale, n_layers, sizes, W, b, act, out, s);
}

}  // namespace lsml
