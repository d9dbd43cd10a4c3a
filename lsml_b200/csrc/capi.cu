// capi.cu -- the C ABI of liblsml_b200.so (declared in include/lsml_b200.h).
//
// (1) the six drop-in symbols of lsml/util/_cutil/masked_gradient.c operating on HOST pointers
//     (stage to the device, run the sm_100a kernel, copy back) and
// (2) the lsml_* device-pointer entry points used by the Python host layer.
//
// There is no CPU fallback anywhere: if no CUDA device is usable the drop-in symbols print the
// error and abort() (they have no error channel, like the reference's MI_CHECK_INDEX build,
// helpers.c:18-33) and the lsml_* functions return a non-zero status.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <mutex>
#include "lsml_internal.h"
#include "band_features.cuh"
#include "../../include/lsml_b200.h"

namespace lsml {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return (int)e;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace lsml

using namespace lsml;

namespace {

[[noreturn]] void die(const char* fn) {
    fprintf(stderr, "liblsml_b200: %s failed: %s\n", fn, lsml_last_error());
    fflush(stderr);
    abort();
}

// Device buffers of the host-pointer drop-in path.  The reference calls these symbols once per
// evolution iteration on arrays of one size (model.py:390), so the buffers are kept between
// calls (four slots per device, grown on demand, freed at process exit by the driver): a call
// costs its copies and its kernel, not four cudaMalloc / cudaFree pairs (~1 ms each at 256^3).
// One call at a time per process (the mutex): the reference is single-threaded.
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int dev = -1;
    int alloc(size_t bytes) {
        int cur = 0;
        if (check_cuda(cudaGetDevice(&cur), "cudaGetDevice")) return 1;
        if (p != nullptr && dev == cur && cap >= bytes) return 0;
        if (p != nullptr) { cudaFree(p); p = nullptr; cap = 0; }
        if (check_cuda(cudaMalloc(&p, bytes ? bytes : 1), "cudaMalloc")) { p = nullptr; return 1; }
        cap = bytes; dev = cur;
        return 0;
    }
};
std::mutex g_host_path_mutex;
DevBuf g_host_buf[4];

// Host-pointer implementation shared by gmag_os{1,2,3}d.  Output buffers are copied IN first
// so that voxels outside the mask keep whatever the caller had there (the reference only
// writes masked voxels; masked_gradient.py:219 zero-fills before the call).
int host_gmag_os(int ndim, const int* shape, double* A, bool* mask, double* nu, double* gmag,
                 const double* dx) {
    Geom g;
    if (make_geom(g, ndim, shape, 1, dx)) return 1;
    const size_t n = (size_t)g.voxels, nb = n * sizeof(double);
    std::lock_guard<std::mutex> lock(g_host_path_mutex);
    DevBuf &dA = g_host_buf[0], &dM = g_host_buf[1], &dNu = g_host_buf[2], &dG = g_host_buf[3];
    if (dA.alloc(nb) || dM.alloc(n) || dNu.alloc(nb) || dG.alloc(nb)) return 2;
    cudaStream_t s = 0;
    LSML_CUDA(cudaMemcpyAsync(dA.p, A, nb, cudaMemcpyHostToDevice, s));
    LSML_CUDA(cudaMemcpyAsync(dM.p, mask, n, cudaMemcpyHostToDevice, s));
    LSML_CUDA(cudaMemcpyAsync(dNu.p, nu, nb, cudaMemcpyHostToDevice, s));
    LSML_CUDA(cudaMemcpyAsync(dG.p, gmag, nb, cudaMemcpyHostToDevice, s));
    int st = gmag_os_dense<double>(g, (const double*)dA.p, (const u8*)dM.p, (const double*)dNu.p,
                                   (double*)dG.p, s);
    if (st) return st;
    LSML_CUDA(cudaMemcpyAsync(gmag, dG.p, nb, cudaMemcpyDeviceToHost, s));
    LSML_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int host_gradient_centered(int ndim, const int* shape, double* A, bool* mask, double** d,
                           double* gmag, const double* dx, int normalize) {
    Geom g;
    if (make_geom(g, ndim, shape, 1, dx)) return 1;
    const size_t n = (size_t)g.voxels, nb = n * sizeof(double);
    std::lock_guard<std::mutex> lock(g_host_path_mutex);
    DevBuf &dA = g_host_buf[0], &dM = g_host_buf[1], &dD = g_host_buf[2], &dG = g_host_buf[3];
    if (dA.alloc(nb) || dM.alloc(n) || dD.alloc(nb * ndim) || dG.alloc(nb)) return 2;
    cudaStream_t s = 0;
    LSML_CUDA(cudaMemcpyAsync(dA.p, A, nb, cudaMemcpyHostToDevice, s));
    LSML_CUDA(cudaMemcpyAsync(dM.p, mask, n, cudaMemcpyHostToDevice, s));
    for (int a = 0; a < ndim; ++a)
        LSML_CUDA(cudaMemcpyAsync((char*)dD.p + a * nb, d[a], nb, cudaMemcpyHostToDevice, s));
    LSML_CUDA(cudaMemcpyAsync(dG.p, gmag, nb, cudaMemcpyHostToDevice, s));
    int st = gradient_centered_dense<double>(g, (const double*)dA.p, (const u8*)dM.p,
                                             (double*)dD.p, (double*)dG.p, normalize, s);
    if (st) return st;
    for (int a = 0; a < ndim; ++a)
        LSML_CUDA(cudaMemcpyAsync(d[a], (char*)dD.p + a * nb, nb, cudaMemcpyDeviceToHost, s));
    LSML_CUDA(cudaMemcpyAsync(gmag, dG.p, nb, cudaMemcpyDeviceToHost, s));
    LSML_CUDA(cudaStreamSynchronize(s));
    return 0;
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------------------------------------
// (1) drop-in symbols
// ---------------------------------------------------------------------------------------------
void gmag_os3d(int m, int n, int p, double* A, bool* mask, double* nu, double* gmag,
               double deli, double delj, double delk) {
    const int shape[3] = {m, n, p};
    const double dx[3] = {deli, delj, delk};
    if (host_gmag_os(3, shape, A, mask, nu, gmag, dx)) die("gmag_os3d");
}
void gmag_os2d(int m, int n, double* A, bool* mask, double* nu, double* gmag, double deli,
               double delj) {
    const int shape[2] = {m, n};
    const double dx[2] = {deli, delj};
    if (host_gmag_os(2, shape, A, mask, nu, gmag, dx)) die("gmag_os2d");
}
void gmag_os1d(int m, double* A, bool* mask, double* nu, double* gmag, double deli) {
    if (host_gmag_os(1, &m, A, mask, nu, gmag, &deli)) die("gmag_os1d");
}
void gradient_centered3d(int m, int n, int p, double* A, bool* mask, double* di, double* dj,
                         double* dk, double* gmag, double deli, double delj, double delk,
                         int normalize) {
    const int shape[3] = {m, n, p};
    const double dx[3] = {deli, delj, delk};
    double* d[3] = {di, dj, dk};
    if (host_gradient_centered(3, shape, A, mask, d, gmag, dx, normalize)) die("gradient_centered3d");
}
void gradient_centered2d(int m, int n, double* A, bool* mask, double* di, double* dj,
                         double* gmag, double deli, double delj, int normalize) {
    const int shape[2] = {m, n};
    const double dx[2] = {deli, delj};
    double* d[2] = {di, dj};
    if (host_gradient_centered(2, shape, A, mask, d, gmag, dx, normalize)) die("gradient_centered2d");
}
void gradient_centered1d(int m, double* A, bool* mask, double* di, double* gmag, double deli,
                         int normalize) {
    double* d[1] = {di};
    if (host_gradient_centered(1, &m, A, mask, d, gmag, &deli, normalize)) die("gradient_centered1d");
}

// ---------------------------------------------------------------------------------------------
// (2) device entry points
// ---------------------------------------------------------------------------------------------
int lsml_version(void) { return 100; }
const char* lsml_last_error(void) { return lsml::g_err; }
long long lsml_launch_count(void) { return lsml::g_launches.load(); }

int lsml_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { check_cuda(e, "cudaGetDeviceCount"); return -(int)e; }
    return n;
}

int lsml_gmag_os_f64(int ndim, const int* shape, long long batch, const double* A,
                     const uint8_t* mask, const double* nu, double* gmag, const double* dx,
                     void* stream) {
    Geom g;
    if (make_geom(g, ndim, shape, batch, dx)) return 1;
    return gmag_os_dense<double>(g, A, mask, nu, gmag, (cudaStream_t)stream);
}
int lsml_gmag_os_f32(int ndim, const int* shape, long long batch, const float* A,
                     const uint8_t* mask, const float* nu, float* gmag, const double* dx,
                     void* stream) {
    Geom g;
    if (make_geom(g, ndim, shape, batch, dx)) return 1;
    return gmag_os_dense<float>(g, A, mask, nu, gmag, (cudaStream_t)stream);
}
int lsml_gradient_centered_f64(int ndim, const int* shape, long long batch, const double* A,
                               const uint8_t* mask, double* grads, double* gmag,
                               const double* dx, int normalize, void* stream) {
    Geom g;
    if (make_geom(g, ndim, shape, batch, dx)) return 1;
    return gradient_centered_dense<double>(g, A, mask, grads, gmag, normalize, (cudaStream_t)stream);
}
int lsml_gradient_centered_f32(int ndim, const int* shape, long long batch, const float* A,
                               const uint8_t* mask, float* grads, float* gmag,
                               const double* dx, int normalize, void* stream) {
    Geom g;
    if (make_geom(g, ndim, shape, batch, dx)) return 1;
    return gradient_centered_dense<float>(g, A, mask, grads, gmag, normalize, (cudaStream_t)stream);
}

long long lsml_compact_workspace_bytes(long long total) { return compact_workspace_bytes(total); }

int lsml_compact_mask(long long total, long long voxels_per_item, const uint8_t* mask,
                      long long* band_idx, long long* n_band, long long* item_offsets_out,
                      void* workspace, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    int st = compact_mask(total, mask, band_idx, n_band, workspace, 0, s);
    if (st) return st;
    if (item_offsets_out != nullptr && voxels_per_item > 0)
        return item_offsets(band_idx, n_band, voxels_per_item, total / voxels_per_item,
                            item_offsets_out, s);
    return 0;
}
int lsml_compact_mask_range(long long total, const uint8_t* mask, long long* band_idx,
                            long long* n_band, long long index_offset, void* workspace,
                            void* stream) {
    return compact_mask(total, mask, band_idx, n_band, workspace, index_offset, (cudaStream_t)stream);
}

#define GEOM(g) Geom g; if (make_geom(g, ndim, shape, batch, dx)) return 1

int lsml_gauss1d_f64(int ndim, const int* shape, long long batch, const double* in, double* out,
                     const double* weights, int radius, int antisym, int axis, int mode,
                     void* stream) {
    const double* dx = nullptr;
    GEOM(g);
    if (axis < 0 || axis >= ndim) { set_error("gauss1d: axis %d out of range", axis); return 1; }
    return gauss1d(g, in, out, weights, radius, antisym, 3 - ndim + axis, mode, (cudaStream_t)stream);
}
int lsml_gauss2d_f64(const int* shape, long long batch, const double* in, double* out,
                     const double* w0, int r0, const double* w1, int r1, int* done, void* stream) {
    const int ndim = 2;
    const double* dx = nullptr;
    GEOM(g);
    bool d = false;
    const int st = gauss2d_small(g, in, out, w0, r0, w1, r1, &d, (cudaStream_t)stream);
    if (done) *done = d ? 1 : 0;
    return st;
}
int lsml_np_gradient_mag_f64(int ndim, const int* shape, long long batch, const double* in,
                             double* out, const double* dx, void* stream) {
    GEOM(g);
    return np_gradient_mag(g, in, out, (cudaStream_t)stream);
}
long long lsml_normalize_workspace_bytes(long long voxels, long long batch) {
    return normalize_workspace_bytes(voxels, batch);
}
int lsml_normalize_f64(long long voxels, long long batch, const double* img, double* out,
                       void* workspace, void* stream) {
    return normalize_images(voxels, batch, img, out, workspace, (cudaStream_t)stream);
}
int lsml_global_stats(int ndim, const int* shape, long long batch, const double* u,
                      unsigned long long* stats, void* stream) {
    const double* dx = nullptr;
    GEOM(g);
    return global_stats(g, u, stats, (cudaStream_t)stream);
}
long long lsml_band_stats_workspace_bytes(long long voxels, long long batch, int nplanes) {
    return band_stats_workspace_bytes(voxels, batch, nplanes);
}
int lsml_band_stats(long long voxels, long long batch, const double* planes,
                    long long plane_stride, int nplanes, const long long* band_idx,
                    const long long* item_offsets, double* mean, double* stdv, void* workspace,
                    void* stream) {
    return band_stats(voxels, batch, planes, plane_stride, nplanes, band_idx, item_offsets, mean,
                      stdv, workspace, (cudaStream_t)stream);
}
long long lsml_contour_workspace_bytes(int m, int n, long long batch) {
    return contour_workspace_bytes(m, n, batch);
}
int lsml_contour_length(int ndim, const int* shape, long long batch, const double* u,
                        const double* dx, double* boundary, void* workspace, void* stream) {
    GEOM(g);
    return contour_length(g, u, boundary, workspace, (cudaStream_t)stream);
}
int lsml_item_features(int nfeat, const int* kind, const int* a, const int* b, int ndim,
                       const int* shape, long long batch, const double* dx, int nplanes,
                       const unsigned long long* stats, const double* band_mean,
                       const double* band_std, const double* boundary, double* item_feat,
                       double* com, void* stream) {
    GEOM(g);
    return item_features_c(nfeat, kind, a, b, g, nplanes, stats, band_mean, band_std, boundary,
                           item_feat, com, (cudaStream_t)stream);
}
int lsml_assemble_features(int nfeat, const int* kind, const int* a, const int* b, int ndim,
                           const int* shape, long long batch, const double* dx,
                           const long long* band_idx, const long long* n_band,
                           const double* planes, long long plane_stride,
                           const double* item_feat, const double* com, double* X, void* stream) {
    GEOM(g);
    return assemble_features_c(nfeat, kind, a, b, g, band_idx, n_band, planes, plane_stride,
                               item_feat, com, X, 0, (cudaStream_t)stream);
}
int lsml_assemble_features_origin(int nfeat, const int* kind, const int* a, const int* b, int ndim,
                                  const int* shape, long long batch, const double* dx,
                                  const long long* band_idx, const long long* n_band,
                                  const double* planes, long long plane_stride,
                                  const double* item_feat, const double* com, double* X,
                                  int origin0, void* stream) {
    GEOM(g);
    return assemble_features_c(nfeat, kind, a, b, g, band_idx, n_band, planes, plane_stride,
                               item_feat, com, X, origin0, (cudaStream_t)stream);
}
int lsml_band_sums(long long voxels, long long batch, const double* planes,
                   long long plane_stride, int nplanes, const long long* band_idx,
                   const long long* item_offsets, int mode, const double* mean, double* sums,
                   void* workspace, void* stream) {
    return band_sums(voxels, batch, planes, plane_stride, nplanes, band_idx, item_offsets, mode,
                     mean, sums, workspace, (cudaStream_t)stream);
}
int lsml_stencil_use_tma(int on) { return stencil_use_tma(on); }

long long lsml_axis_moments_workspace_bytes(int ndim, const int* shape, long long batch) {
    const double* dx = nullptr;
    Geom g;
    if (make_geom(g, ndim, shape, batch, dx)) return -1;
    return axis_moments_workspace_bytes(g);
}
int lsml_axis_moments(int ndim, const int* shape, long long batch, const double* dx,
                      const double* u, const unsigned long long* stats, const double* com,
                      int ncols, const int* cols, const int* axes, const int* orders, int nfeat,
                      double* item_feat, void* workspace, void* stream) {
    GEOM(g);
    return axis_moments(g, u, stats, com, ncols, cols, axes, orders, nfeat, item_feat, workspace,
                        (cudaStream_t)stream);
}
long long lsml_surface_area_workspace_bytes(int ndim, const int* shape, long long batch) {
    const double* dx = nullptr;
    Geom g;
    if (make_geom(g, ndim, shape, batch, dx)) return -1;
    return surface_area_workspace_bytes(g);
}
int lsml_surface_area(int ndim, const int* shape, long long batch, const double* u,
                      const double* dx, double* area, void* workspace, void* stream) {
    GEOM(g);
    return surface_area(g, u, area, workspace, (cudaStream_t)stream);
}
int lsml_linear_predict(const double* X, const long long* n_band, int nfeat, const double* mean,
                        const double* scale, const double* coef, double intercept, double* out,
                        void* stream) {
    return linear_predict(X, n_band, nfeat, mean, scale, coef, intercept, out, (cudaStream_t)stream);
}
int lsml_forest_predict(const double* X, const long long* n_band, int nfeat, const double* mean,
                        const double* scale, const void* nodes, const double* leaves,
                        const unsigned* root_ref, const uint8_t* root_is_leaf, int n_trees,
                        double divisor, const void* nodes8, const int* tree_base, int max_nodes,
                        double* out, void* stream) {
    return forest_predict(X, n_band, nfeat, mean, scale, nodes, leaves, root_ref, root_is_leaf,
                          n_trees, divisor, nodes8, tree_base, max_nodes, out, (cudaStream_t)stream);
}
int lsml_mlp_predict(const double* X, const long long* n_band, const double* mean,
                     const double* scale, int n_layers, const int* sizes,
                     const double* const* W, const double* const* b, int act, double* out,
                     void* stream) {
    return mlp_predict(X, n_band, mean, scale, n_layers, sizes, W, b, act, out, (cudaStream_t)stream);
}

// the feature program as the source of the regressor's inputs (no materialised matrix)
static int band_source(const lsml_band_source* src, BandFeatures& bf) {
    if (src == nullptr) { set_error("band source is NULL"); return 1; }
    const int ndim = src->ndim;
    const int* shape = src->shape;
    const long long batch = src->batch;
    const double* dx = src->dx;
    GEOM(g);
    FeatProgram prog;
    if (fill_program(prog, src->nfeat, src->kind, src->a, src->b)) return 1;
    for (int f = 0; f < src->nfeat; ++f)
        if (src->kind[f] == 8 /* FK_COM_RAY */ && (src->origin0 != 0 || ndim < 2)) {
            set_error("COMRaySample needs the whole 2-D/3-D image on one device");
            return 1;
        }
    bf = make_band_features(prog, g, src->band_idx, src->planes, src->plane_stride,
                            src->item_feat, src->com, src->origin0);
    return 0;
}
int lsml_linear_predict_band(const lsml_band_source* src, const long long* n_band,
                             const double* mean, const double* scale, const double* coef,
                             double intercept, double* out, void* stream) {
    BandFeatures bf;
    if (band_source(src, bf)) return 1;
    return linear_predict_band(bf, n_band, mean, scale, coef, intercept, out, (cudaStream_t)stream);
}
int lsml_forest_predict_band(const lsml_band_source* src, const long long* n_band,
                             const double* mean, const double* scale, const void* nodes,
                             const double* leaves, const unsigned* root_ref,
                             const uint8_t* root_is_leaf, int n_trees, double divisor,
                             const void* nodes8, const int* tree_base, int max_nodes, double* out,
                             void* stream) {
    BandFeatures bf;
    if (band_source(src, bf)) return 1;
    return forest_predict_band(bf, n_band, mean, scale, nodes, leaves, root_ref, root_is_leaf,
                               n_trees, divisor, nodes8, tree_base, max_nodes, out,
                               (cudaStream_t)stream);
}
int lsml_mlp_predict_band(const lsml_band_source* src, const long long* n_band, const double* mean,
                          const double* scale, int n_layers, const int* sizes,
                          const double* const* W, const double* const* b, int act, double* out,
                          void* stream) {
    BandFeatures bf;
    if (band_source(src, bf)) return 1;
    return mlp_predict_band(bf, n_band, mean, scale, n_layers, sizes, W, b, act, out,
                            (cudaStream_t)stream);
}
static int ranked_args(const lsml_ranked_forest* f, RankedArgs& a) {
    if (f == nullptr) { set_error("ranked forest is NULL"); return 1; }
    a = RankedArgs{f->words, f->tree_start, f->root_is_leaf, f->thr_tab, f->thr_off, f->n_trees,
                   f->nfeat, f->fbits, f->rbits, f->max_tree_words, f->divisor};
    return 0;
}
int lsml_forest_ranked_predict(const double* X, const long long* n_band, int nfeat,
                               const double* mean, const double* scale,
                               const lsml_ranked_forest* forest, double* out, void* stream) {
    RankedArgs a;
    if (ranked_args(forest, a)) return 1;
    return forest_ranked_predict(X, n_band, nfeat, mean, scale, a, out, (cudaStream_t)stream);
}
int lsml_forest_ranked_predict_band(const lsml_band_source* src, const long long* n_band,
                                    const double* mean, const double* scale,
                                    const lsml_ranked_forest* forest, double* out, void* stream) {
    BandFeatures bf;
    RankedArgs a;
    if (band_source(src, bf) || ranked_args(forest, a)) return 1;
    return forest_ranked_predict_band(bf, n_band, mean, scale, a, out, (cudaStream_t)stream);
}
int lsml_band_update(int ndim, const int* shape, long long batch, const double* dx, double* u,
                     const double* vel, const long long* band_idx, const long long* n_band,
                     double step, double* new_u, double* gmag_out, int apply, void* stream) {
    GEOM(g);
    return band_update(g, u, vel, band_idx, n_band, step, new_u, gmag_out, apply, nullptr,
                       (cudaStream_t)stream);
}
int lsml_band_update_stats(int ndim, const int* shape, long long batch, const double* dx, double* u,
                           const double* vel, const long long* band_idx, const long long* n_band,
                           double step, double* new_u, double* gmag_out,
                           unsigned long long* stats, void* stream) {
    GEOM(g);
    return band_update(g, u, vel, band_idx, n_band, step, new_u, gmag_out, 1, stats,
                       (cudaStream_t)stream);
}
long long lsml_fmm_workspace_bytes(long long total) { return fmm_workspace_bytes(total); }
int lsml_fmm_workspace_init(long long total, void* workspace, void* stream) {
    return fmm_workspace_init(total, workspace, (cudaStream_t)stream);
}
int lsml_fmm_rebuild(int ndim, const int* shape, long long batch, const double* dx,
                     const double* u, double band, const unsigned long long* stats,
                     uint8_t* mask, double* dist_out, void* workspace, void* stream) {
    GEOM(g);
    return fmm_rebuild(g, u, band, stats, mask, dist_out, workspace, nullptr, nullptr,
                       (cudaStream_t)stream);
}
int lsml_fmm_rebuild_from_band(int ndim, const int* shape, long long batch, const double* dx,
                               const double* u, double band, const unsigned long long* stats,
                               uint8_t* mask, double* dist_out, void* workspace,
                               const long long* prev_band_idx, const long long* prev_n_band,
                               void* stream) {
    GEOM(g);
    return fmm_rebuild(g, u, band, stats, mask, dist_out, workspace, prev_band_idx, prev_n_band,
                       (cudaStream_t)stream);
}
int lsml_fmm_slab_begin(int ndim, const int* shape, const double* dx, int own_lo, int own_hi,
                        const double* u, uint8_t* mask, void* workspace,
                        const long long* prev_band_idx, const long long* prev_n_band,
                        const int* dense_planes, void* stream) {
    const long long batch = 1;
    GEOM(g);
    return fmm_slab_begin(g, own_lo, own_hi, u, mask, workspace, prev_band_idx, prev_n_band,
                          dense_planes, (cudaStream_t)stream);
}
int lsml_fmm_slab_march(int ndim, const int* shape, const double* dx, int own_lo, int own_hi,
                        const double* u, double band, uint8_t* mask, void* workspace,
                        void* stream) {
    const long long batch = 1;
    GEOM(g);
    return fmm_slab_march(g, own_lo, own_hi, u, band, mask, workspace, (cudaStream_t)stream);
}
int lsml_fmm_slab_requeue(int ndim, const int* shape, const double* dx, int own_lo, int own_hi,
                          double band, const long long* sides_dev, uint8_t* mask,
                          void* workspace, long long* pending_out, void* stream) {
    const long long batch = 1;
    GEOM(g);
    return fmm_slab_requeue(g, own_lo, own_hi, band, sides_dev, mask, workspace, pending_out,
                            (cudaStream_t)stream);
}
int lsml_fmm_slab_finish(int ndim, const int* shape, const double* dx, int own_lo, int own_hi,
                         double band, uint8_t* mask, double* dist_out, void* workspace,
                         void* stream) {
    const long long batch = 1;
    GEOM(g);
    return fmm_slab_finish(g, own_lo, own_hi, band, mask, dist_out, workspace,
                           (cudaStream_t)stream);
}
int lsml_fmm_counters(long long total, void* workspace, long long* out, void* stream) {
    const i64* dev = nullptr;
    fmm_counters_ptr(total, workspace, &dev);
    LSML_CUDA(cudaMemcpyAsync(out, dev, 8 * sizeof(i64), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    LSML_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

int lsml_fmm_sweep_sizes(long long total, void* workspace, long long* out, void* stream) {
    const i64* dev = nullptr;
    fmm_counters_ptr(total, workspace, &dev);
    LSML_CUDA(cudaMemcpyAsync(out, reinterpret_cast<const char*>(dev) + fmm_counters_sizes_offset(),
                              64 * sizeof(i64), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    LSML_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

long long lsml_fmm_counters_offset(long long total) { return fmm_counters_offset(total); }
long long lsml_fmm_status_offset(void) { return fmm_counters_failed_offset(); }
long long lsml_fmm_phase_offset(void) { return fmm_counters_phase_offset(); }

int lsml_ball_init(int ndim, const int* shape, long long batch, const double* dx, double* u,
                   const double* location, double radius, const double* centres,
                   const double* radii, void* stream) {
    GEOM(g);
    return ball_init(g, u, location, radius, centres, radii, (cudaStream_t)stream);
}
int lsml_overlap_counts(long long voxels, long long batch, const double* u, const uint8_t* seg,
                        unsigned long long* counts, void* stream) {
    return overlap_counts(voxels, batch, u, seg, counts, (cudaStream_t)stream);
}

}  // extern "C"
