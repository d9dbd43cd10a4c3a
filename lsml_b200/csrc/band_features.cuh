// band_features.cuh -- the band feature columns as a device function.
//
// A feature column is one entry (kind, a, b) of the program lsml_b200/core/engine.py compiles
// from the feature list (lsml/feature/feature_map.py:42-51: column order = feature list order).
// `band_feature()` evaluates column f for one band voxel; it is the single definition used by
//   * assemble_kernel (features.cu), which materialises the (B, F) matrix for FeatureMap.__call__
//     and for `fit` (the regression matrix of fit_job_handler.py:337-377), and
//   * the fused velocity kernels (regress.cu), which evaluate the columns in registers / shared
//     memory and never write the matrix ("algorithmic minimum: not materialised").
// The includer provides i64 / u64 (lsml_internal.h, or tests/native/cuda_host_shim.h when
// tests/test_feature_kernels_host.py compiles the marked region with g++).
#pragma once
#include "com_ray.cuh"

namespace lsml {

// [host-test:begin] (tests/test_feature_kernels_host.py compiles this region with g++)
// ---------------------------------------------------------------------------------------------
// per-item ("global") feature values
// ---------------------------------------------------------------------------------------------
enum FeatKind {
    FK_PLANE = 0,       // a = plane index (local)
    FK_DIST_COM = 1,    // local
    FK_BAND_MEAN = 2,   // a = plane index
    FK_BAND_STD = 3,    // a = plane index
    FK_SIZE = 4,
    FK_BOUNDARY = 5,
    FK_ISOPERIMETRIC = 6,
    FK_MOMENT = 7,      // a = axis (0..ndim-1), b = order (1 or 2)
    FK_COM_RAY = 8,     // a = plane index of the raw image, b = sample j | (n_samples << 16) (local)
};

struct FeatProgram {
    int nfeat;
    int kind[64];
    int a[64];
    int b[64];
};

struct ItemGeom {
    int ndim;
    int pad;           // 3 - ndim
    double dx[3];      // padded
    double pdx;        // prod(dx) over real axes, multiplied left to right
    int origin0;       // index of the local grid's first axis-0 plane in the whole volume (slabs)
    int n0;            // extent of the (padded) first axis
};

// moment of `order` along padded axis ax from the exact integer sums
__device__ __forceinline__ double moment_from_stats(const u64* st, int ax, int order, double dx, double pdx) {
    const double cnt = (double)st[0];
    const double measure = cnt * pdx;                       // Size: (u>0).sum()*prod(dx)
    if (order == 1) return ((double)st[1 + ax] * dx) * pdx / measure;   // shape.py:227-228
    if (order > 2) return 0.0;     // filled in afterwards by hi_moment_kernel (features.cu)
    // order 2: sum (x - c)^2 with c the centroid = (N*Sxx - Sx^2)/N exactly (128-bit integers)
    const unsigned __int128 N = st[0], Sx = st[1 + ax], Sxx = st[4 + ax];
    const unsigned __int128 num = N * Sxx - Sx * Sx;
    const double numd = (double)(u64)(num >> 64) * 18446744073709551616.0 + (double)(u64)num;
    const double ssd = (numd / cnt) * (dx * dx);
    return ssd * pdx / measure;
}


// Where the columns of a band voxel come from (all DEVICE pointers; lsml_b200.h: lsml_band_source)
struct BandFeatures {
    FeatProgram prog;
    ItemGeom geo;
    int n1, n2;
    i64 voxels;
    const i64* band_idx;
    const double* planes;      // (P, batch * voxels)
    i64 plane_stride;
    const double* item_feat;   // [item][nfeat]
    const double* com;         // [item][3] padded axes
};

// a band voxel: flat index, item, local coordinates (decoded once, used by every column)
struct VoxelRef {
    i64 g, item;
    int i, j, k;
};

__device__ __forceinline__ VoxelRef voxel_ref(const BandFeatures& s, i64 g) {
    VoxelRef v;
    v.g = g;
    v.item = g / s.voxels;
    const i64 loc = g - v.item * s.voxels;
    v.k = (int)(loc % s.n2);
    const i64 r = loc / s.n2;
    v.j = (int)(r % s.n1);
    v.i = (int)(r / s.n1);
    return v;
}

// value of column f at band voxel v (feature_map.py:53-80 restricted to one voxel)
__device__ __forceinline__ double band_feature(const BandFeatures& s, const VoxelRef& v, int f) {
    const int kind = s.prog.kind[f];
    if (kind == FK_PLANE) return __ldg(s.planes + (i64)s.prog.a[f] * s.plane_stride + v.g);
    if (kind == FK_DIST_COM) {
        const double* c = s.com + v.item * 3;
        // numpy.linalg.norm(mesh - com, axis=0): ((d0^2 + d1^2) + d2^2) over the real axes
        double acc = 0.0;
        if (s.geo.ndim == 3) { const double d = (v.i + s.geo.origin0) * s.geo.dx[0] - c[0]; acc = d * d; }
        if (s.geo.ndim >= 2) { const double d = v.j * s.geo.dx[1] - c[1]; acc = s.geo.ndim == 2 ? d * d : acc + d * d; }
        { const double d = v.k * s.geo.dx[2] - c[2]; acc = s.geo.ndim == 1 ? d * d : acc + d * d; }
        return sqrt(acc);
    }
    if (kind == FK_COM_RAY) {
        // COMRaySample (image.py:147-171): the raw image along the voxel -> centroid ray
        const int idx[3] = {v.i + s.geo.origin0, v.j, v.k};
        const int ext[3] = {s.geo.n0, s.n1, s.n2};
        return com_ray_sample(s.planes + (i64)s.prog.a[f] * s.plane_stride + v.item * s.voxels,
                              s.geo.ndim, ext, s.geo.dx, s.com + v.item * 3, idx,
                              s.prog.b[f] & 0xffff, s.prog.b[f] >> 16);
    }
    return s.item_feat[v.item * s.prog.nfeat + f];
}
// [host-test:end]

}  // namespace lsml
