// Host build of the product's COMRaySample arithmetic (lsml_b200/csrc/com_ray.cuh is
// __host__ __device__): lets tests/test_com_ray_host.py check the kernel's per-sample function
// against the reference's golden vectors without a GPU.  Test infrastructure only.
#include "com_ray.cuh"

extern "C" void com_ray_host(const double* img, int ndim, const int* n, const double* dx,
                             const double* com, const long long* coords, long long n_band,
                             int n_samples, double* out) {
    for (long long b = 0; b < n_band; ++b) {
        int idx[3];
        for (int a = 0; a < 3; ++a) idx[a] = (int)coords[b * 3 + a];
        for (int j = 0; j < 2 * n_samples; ++j)
            out[b * 2 * n_samples + j] = lsml::com_ray_sample(img, ndim, n, dx, com, idx, j, n_samples);
    }
}
