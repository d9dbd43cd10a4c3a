// com_ray.cuh -- one sample of COMRaySample (lsml/feature/provided/image.py:124-171).
//
// For a band voxel `coord` and the centre of mass `com` of {u > 0} the reference evaluates the
// RAW image (sigma only appears in the feature's name) at
//     points = t * com + (1 - t) * (coord * dx),   t = numpy.linspace(-1, 1, 2N+2)[1:-1]
// with scipy.interpolate.RegularGridInterpolator(method='linear', bounds_error=False,
// fill_value=0.0) on the grid arange(shape[a]) * dx[a].  This header restates that arithmetic
// operation by operation (numpy.linspace: arange * step + start; scipy find_indices: interval
// with grid[i] <= x < grid[i+1], clipped to [0, n-2]; 2-D: scipy's evaluate_linear_2d order
// ((v * w0) * w1); 3-D: RegularGridInterpolator._evaluate_linear, weight = (w0 * w1) * w2, corners
// in itertools.product order), so that with -fmad=false the result is bit-identical.
//
// The function is __host__ __device__ on purpose: tests/test_com_ray_host.py compiles it with
// g++ and checks it against the reference's golden vectors without a GPU.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define LSML_HD __host__ __device__ __forceinline__
#else
#define LSML_HD inline
#endif

namespace lsml {

struct RayCell {
    long long i;   // lower grid index of the interval
    double y;      // normalised distance to it
    bool outside;  // x < grid[0] or x > grid[n-1]
};

// grid[i] = (double)i * delta, as numpy.arange(n, dtype=float) * delta builds it
LSML_HD RayCell ray_cell(double x, double delta, int n) {
    RayCell c;
    c.outside = (x < 0.0 * delta) || (x > (double)(n - 1) * delta);
    const double q = x / delta;
    long long i = q < 0.0 ? 0 : (q > (double)(n - 2) ? (long long)(n - 2) : (long long)q);
    while (i > 0 && x < (double)i * delta) --i;
    while (i < n - 2 && x >= (double)(i + 1) * delta) ++i;
    const double lo = (double)i * delta, hi = (double)(i + 1) * delta;
    c.i = i;
    c.y = (x - lo) / (hi - lo);
    return c;
}

// img: one item, C order over the padded extents n[3]; dx, com, idx padded alike (axis a of an
// ndim grid sits at 3-ndim+a).  j = sample index 0 .. 2*n_samples-1.
LSML_HD double com_ray_sample(const double* img, int ndim, const int* n, const double* dx,
                              const double* com, const int* idx, int j, int n_samples) {
    const int pad = 3 - ndim;
    const double step = 2.0 / (double)(2 * n_samples + 1);
    const double t = (double)(j + 1) * step + (-1.0);
    const double s = 1.0 - t;
    RayCell c[3];
    bool outside = false;
    double bad = 0.0;      // the NaN coordinate, if any (scipy: f(nan) = nan, after the fill)
    for (int a = pad; a < 3; ++a) {
        const double x = t * com[a] + s * ((double)idx[a] * dx[a]);
        if (x != x) { bad = x; c[a].i = 0; c[a].y = 0.0; c[a].outside = false; continue; }
        c[a] = ray_cell(x, dx[a], n[a]);
        outside = outside || c[a].outside;
    }
    if (bad != bad) return bad;
    if (outside) return 0.0;
    const long long s1 = n[2], s0 = (long long)n[1] * n[2];
    if (ndim == 2) {
        const double* p = img + c[1].i * s1 + c[2].i;
        const double y0 = c[1].y, y1 = c[2].y;
        double v = 0.0;
        v = v + p[0] * (1 - y0) * (1 - y1);
        v = v + p[1] * (1 - y0) * y1;
        v = v + p[s1] * y0 * (1 - y1);
        v = v + p[s1 + 1] * y0 * y1;
        return v;
    }
    const double* p = img + c[0].i * s0 + c[1].i * s1 + c[2].i;
    const double w0[2] = {1 - c[0].y, c[0].y};
    const double w1[2] = {1 - c[1].y, c[1].y};
    const double w2[2] = {1 - c[2].y, c[2].y};
    double v = 0.0;
    for (int a0 = 0; a0 < 2; ++a0)
        for (int a1 = 0; a1 < 2; ++a1)
            for (int a2 = 0; a2 < 2; ++a2) {
                const double w = ((1.0 * w0[a0]) * w1[a1]) * w2[a2];
                v = v + p[a0 * s0 + a1 * s1 + a2] * w;
            }
    return v;
}

}  // namespace lsml
