// regress.cu -- inference of the host-trained regressor on the band feature matrix
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
// Forest layout: 8-byte nodes, children adjacent (right = left + 1):
//   internal  { float thr; u32 meta = feature<<24 | left_is_leaf<<23 | right_is_leaf<<22 | left }
//   leaf      { double value }            (the parent's meta says which interpretation applies)
// A 100-tree forest of ~600-node trees is ~480 KB: L2-resident, partly L1-resident.  The
// traversal is latency-bound pointer chasing (not HBM-bound); one thread per band voxel with
// the voxel's float32 features staged in shared memory as [feature][thread] (conflict-free).
// Measured alternative (round 1, tools/forest_probe.py, 270k x 16, RF-100): walking the forest
// tree by tree with the current tree staged in shared memory and 4 voxels per thread was 2x
// SLOWER (738 us vs 388 us): the plain kernel runs at full occupancy (64 warps/SM) and is bound
// by L1 tag throughput on divergent node fetches, which the staged version only replaces by
// shared-memory bank conflicts at a quarter of the occupancy.  Walking 2 or 4 trees per thread
// at once (independent chains for latency hiding) was also slower (578 / 543 us): the limit is
// the number of divergent L1 sectors per warp-load, not the latency of a single chain.
#include "lsml_internal.h"

namespace lsml {

struct Scaler {
    const double* mean;   // may be null => no scaler
    const double* scale;
};

__device__ __forceinline__ double scaled(const double* __restrict__ X, i64 row, int f, int F,
                                         const Scaler& sc) {
    double x = __ldg(X + row * F + f);
    if (sc.mean != nullptr) x = (x - __ldg(sc.mean + f)) / __ldg(sc.scale + f);
    return x;
}

// --------------------------------------------------------------------------------- linear
__global__ void __launch_bounds__(256)
linear_predict_kernel(const double* __restrict__ X, const i64* __restrict__ n_band, int F,
                      Scaler sc, const double* __restrict__ coef, double intercept,
                      double* __restrict__ out) {
    const i64 nb = *n_band;
    for (i64 b = (i64)blockIdx.x * blockDim.x + threadIdx.x; b < nb;
         b += (i64)gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int f = 0; f < F; ++f) acc += scaled(X, b, f, F, sc) * __ldg(coef + f);
        out[b] = acc + intercept;
    }
}

// --------------------------------------------------------------------------------- forest
struct ForestView {
    const uint2* nodes;        // 8-byte nodes of all trees back to back
    const int* tree_base;      // [n_trees + 1] first node of each tree, then the node count
    const u8* root_is_leaf;    // [n_trees]
    int n_trees;
    double divisor;            // n_estimators for forests, 1 for a single tree
};

constexpr int FOREST_THREADS = 128;

__global__ void __launch_bounds__(FOREST_THREADS)
forest_predict_kernel(const double* __restrict__ X, const i64* __restrict__ n_band, int F,
                      Scaler sc, ForestView fv, double* __restrict__ out) {
    extern __shared__ float feat[];   // [F][FOREST_THREADS]
    const i64 nb = *n_band;
    const i64 nblocks = (nb + FOREST_THREADS - 1) / FOREST_THREADS;
    for (i64 blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const i64 b = blk * FOREST_THREADS + threadIdx.x;
        const bool live = b < nb;
        __syncthreads();
        if (live)
            for (int f = 0; f < F; ++f)
                feat[f * FOREST_THREADS + threadIdx.x] = __double2float_rn(scaled(X, b, f, F, sc));
        __syncthreads();
        if (!live) continue;
        double acc = 0.0;
        for (int t = 0; t < fv.n_trees; ++t) {
            const uint2* tree = fv.nodes + __ldg(fv.tree_base + t);
            int idx = 0;
            bool leaf = __ldg(fv.root_is_leaf + t) != 0;
            while (!leaf) {
                const uint2 nd = __ldg(tree + idx);
                const float thr = __uint_as_float(nd.x);
                const unsigned meta = nd.y;
                const float x = feat[(meta >> 24) * FOREST_THREADS + threadIdx.x];
                const bool right = !(x <= thr);
                idx = (int)(meta & 0x3fffffu) + (right ? 1 : 0);
                leaf = right ? ((meta >> 22) & 1u) : ((meta >> 23) & 1u);
            }
            const uint2 lv = __ldg(tree + idx);
            acc += __hiloint2double((int)lv.y, (int)lv.x);
        }
        out[b] = acc / fv.divisor;
    }
}

// --------------------------------------------------------------------------------- MLP
// One thread per band voxel; weights are read through the read-only path (L1-resident: a
// (F x H) float64 layer with F=16, H=100 is 12.8 KB).  Activations ping-pong in shared memory
// as [unit][thread].  act: 0 identity, 1 relu, 2 tanh, 3 logistic.
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
mlp_predict_kernel(const double* __restrict__ X, const i64* __restrict__ n_band, Scaler sc,
                   MlpView mv, double* __restrict__ out) {
    extern __shared__ double actbuf[];   // 2 x [max_width][MLP_THREADS]
    double* cur = actbuf;
    double* nxt = actbuf + (size_t)mv.max_width * MLP_THREADS;
    const i64 nb = *n_band;
    const int F = mv.sizes[0];
    for (i64 b = (i64)blockIdx.x * MLP_THREADS + threadIdx.x; b < nb;
         b += (i64)gridDim.x * MLP_THREADS) {
        for (int f = 0; f < F; ++f) cur[f * MLP_THREADS + threadIdx.x] = scaled(X, b, f, F, sc);
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

// --------------------------------------------------------------------------------- launchers
int linear_predict(const double* X, const i64* n_band, int F, const double* mean,
                   const double* scale, const double* coef, double intercept, double* out,
                   cudaStream_t s) {
    Scaler sc{mean, scale};
    linear_predict_kernel<<<LSML_SMS * 8, 256, 0, s>>>(X, n_band, F, sc, coef, intercept, out);
    LSML_LAUNCH_CHECK("linear_predict_kernel");
    return 0;
}

int forest_predict(const double* X, const i64* n_band, int F, const double* mean,
                   const double* scale, const void* nodes, const int* tree_base,
                   const u8* root_is_leaf, int n_trees, double divisor, double* out,
                   cudaStream_t s) {
    if (F > 255) { set_error("forest: at most 255 features"); return 1; }
    Scaler sc{mean, scale};
    ForestView fv{static_cast<const uint2*>(nodes), tree_base, root_is_leaf, n_trees, divisor};
    const size_t smem = (size_t)F * FOREST_THREADS * sizeof(float);
    if (smem > 48 * 1024)
        LSML_CUDA(cudaFuncSetAttribute(forest_predict_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    forest_predict_kernel<<<LSML_SMS * 16, FOREST_THREADS, smem, s>>>(X, n_band, F, sc, fv, out);
    LSML_LAUNCH_CHECK("forest_predict_kernel");
    return 0;
}

int mlp_predict(const double* X, const i64* n_band, const double* mean, const double* scale,
                int n_layers, const int* sizes, const double* const* W, const double* const* b,
                int act, double* out, cudaStream_t s) {
    if (n_layers < 1 || n_layers > 8) { set_error("mlp: 1..8 layers supported"); return 1; }
    Scaler sc{mean, scale};
    MlpView mv;
    mv.n_layers = n_layers; mv.act = act; mv.max_width = 0;
    for (int l = 0; l <= n_layers; ++l) {
        mv.sizes[l] = sizes[l];
        if (sizes[l] > mv.max_width) mv.max_width = sizes[l];
    }
    for (int l = 0; l < n_layers; ++l) { mv.W[l] = W[l]; mv.b[l] = b[l]; }
    const size_t smem = 2 * (size_t)mv.max_width * MLP_THREADS * sizeof(double);
    if (smem > 200 * 1024) { set_error("mlp: layer width %d too large", mv.max_width); return 1; }
    if (smem > 48 * 1024)
        LSML_CUDA(cudaFuncSetAttribute(mlp_predict_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mlp_predict_kernel<<<LSML_SMS * 4, MLP_THREADS, smem, s>>>(X, n_band, sc, mv, out);
    LSML_LAUNCH_CHECK("mlp_predict_kernel");
    return 0;
}

}  // namespace lsml
