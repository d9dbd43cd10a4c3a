// misc.cu -- small kernels either side of the hot path.
//
//   ball_init_kernel   BallInitializer.initialize + InitializerBase (u0 = 2*mask - 1):
//                      lsml/initializer/provided/ball.py:13-31, initializer_base.py:56
//   overlap_kernel     |{u>0} & seg| and |{u>0} | seg| per item (jaccard / dice,
//                      lsml/score_functions.py:4-30) as exact integer counts
#include "lsml_internal.h"

namespace lsml {

struct BallParams {
    int ndim, n0, n1, n2;
    i64 voxels;
    double dx[3];      // padded
    double cdx[3];     // (location * dx) per padded axis
    double radius;
};

// mask = (radius - sqrt(sum_a (idx_a*dx_a - loc_a*dx_a)^2)) > 0 ; u0 = 2*mask - 1
// centres: per-item centre, centres[item*3 + padded axis], already multiplied by dx
__global__ void __launch_bounds__(256)
ball_init_kernel(double* __restrict__ u, BallParams p, const double* __restrict__ centres,
                 const double* __restrict__ radii, i64 batch) {
    const i64 total = p.voxels * batch;
    for (i64 l = (i64)blockIdx.x * blockDim.x + threadIdx.x; l < total;
         l += (i64)gridDim.x * blockDim.x) {
        const i64 item = l / p.voxels, loc = l - item * p.voxels;
        const int k = (int)(loc % p.n2);
        const i64 r = loc / p.n2;
        const int j = (int)(r % p.n1), i = (int)(r / p.n1);
        const double* c = centres ? centres + item * 3 : p.cdx;
        const double rad = radii ? radii[item] : p.radius;
        double acc = 0.0;
        if (p.ndim == 3) { const double d = i * p.dx[0] - c[0]; acc = d * d; }
        if (p.ndim >= 2) { const double d = j * p.dx[1] - c[1]; acc = p.ndim == 2 ? d * d : acc + d * d; }
        { const double d = k * p.dx[2] - c[2]; acc = p.ndim == 1 ? d * d : acc + d * d; }
        u[l] = (rad - sqrt(acc)) > 0 ? 1.0 : -1.0;
    }
}

__global__ void __launch_bounds__(256)
overlap_kernel(const double* __restrict__ u, const u8* __restrict__ seg, i64 voxels,
               u64* __restrict__ counts /* [item][2] = {intersection, union} */, i64 item0) {
    __shared__ u64 red[32];
    const i64 item = item0 + blockIdx.y;
    const i64 base = (i64)item * voxels;
    u64 inter = 0, uni = 0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < voxels;
         i += (i64)gridDim.x * blockDim.x) {
        const bool a = __ldg(u + base + i) > 0.0, b = __ldg(seg + base + i) != 0;
        inter += (a && b) ? 1 : 0;
        uni += (a || b) ? 1 : 0;
    }
    const u64 ti = block_sum<u64>(inter, red);
    const u64 tu = block_sum<u64>(uni, red);
    if (threadIdx.x == 0) {
        if (ti) atomicAdd(counts + item * 2, ti);
        if (tu) atomicAdd(counts + item * 2 + 1, tu);
    }
}

int ball_init(const Geom& g, double* u, const double* location, double radius,
              const double* centres_dev, const double* radii_dev, cudaStream_t s) {
    BallParams p;
    p.ndim = g.ndim; p.n0 = g.n[0]; p.n1 = g.n[1]; p.n2 = g.n[2]; p.voxels = g.voxels;
    const int pad = 3 - g.ndim;
    for (int a = 0; a < 3; ++a) {
        p.dx[a] = g.dx[a];
        p.cdx[a] = (a >= pad && location) ? location[a - pad] * g.dx[a] : 0.0;
    }
    p.radius = radius;
    ball_init_kernel<<<LSML_SMS * 8, 256, 0, s>>>(u, p, centres_dev, radii_dev, g.batch);
    LSML_LAUNCH_CHECK("ball_init_kernel");
    return 0;
}

int overlap_counts(i64 voxels, i64 batch, const double* u, const u8* seg, u64* counts,
                   cudaStream_t s) {
    LSML_CUDA(cudaMemsetAsync(counts, 0, sizeof(u64) * 2 * batch, s));
    i64 bx = (voxels + 256 * 8 - 1) / (256 * 8);
    if (bx < 1) bx = 1;
    if (bx > 1024) bx = 1024;
    for (i64 item0 = 0; item0 < batch; item0 += 65535) {          // gridDim.y <= 65535
        const i64 ny = batch - item0 < 65535 ? batch - item0 : 65535;
        overlap_kernel<<<dim3((unsigned)bx, (unsigned)ny), 256, 0, s>>>(u, seg, voxels, counts, item0);
    }
    LSML_LAUNCH_CHECK("overlap_kernel");
    return 0;
}

}  // namespace lsml
