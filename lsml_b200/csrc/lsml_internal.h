// lsml_internal.h -- C++ declarations shared between the kernel translation units and capi.cu
#pragma once
#include "common.cuh"

namespace lsml {

// stencil.cu
template <typename T>
int gmag_os_dense(const Geom& g, const T* A, const u8* mask, const T* nu, T* out, cudaStream_t s);
template <typename T>
int gradient_centered_dense(const Geom& g, const T* A, const u8* mask, T* grads, T* gmag,
                            int normalize, cudaStream_t s);

int stencil_use_tma(int on);

// compact.cu
i64 compact_workspace_bytes(i64 total);
int compact_mask(i64 total, const u8* mask, i64* band_idx, i64* n_band, void* workspace,
                 i64 index_offset, cudaStream_t s);
int item_offsets(const i64* band_idx, const i64* n_band, i64 voxels, i64 batch, i64* offsets,
                 cudaStream_t s);

}  // namespace lsml

namespace lsml {

// features.cu
struct FeatProgram;
int gauss1d(const Geom& g, const double* in, double* out, const double* w, int radius,
            int antisym, int axis_padded, int mode, cudaStream_t s);
int gauss2d_small(const Geom& g, const double* in, double* out, const double* w0, int r0,
                  const double* w1, int r1, bool* done, cudaStream_t s);
int np_gradient_mag(const Geom& g, const double* in, double* out, cudaStream_t s);
i64 normalize_workspace_bytes(i64 voxels, i64 batch);
int normalize_images(i64 voxels, i64 batch, const double* img, double* out, void* workspace,
                     cudaStream_t s);
int global_stats(const Geom& g, const double* u, u64* stats, cudaStream_t s);
i64 band_stats_workspace_bytes(i64 voxels, i64 batch, int nplanes);
int band_stats(i64 voxels, i64 batch, const double* planes, i64 plane_stride, int nplanes,
               const i64* band_idx, const i64* item_offsets, double* mean, double* stdv,
               void* workspace, cudaStream_t s);
i64 contour_workspace_bytes(int m, int n, i64 batch);
int contour_length(const Geom& g, const double* u, double* boundary, void* workspace,
                   cudaStream_t s);
int item_features_c(int nfeat, const int* kind, const int* a, const int* b, const Geom& g,
                    int nplanes, const u64* stats, const double* band_mean,
                    const double* band_std, const double* boundary, double* item_feat,
                    double* com, cudaStream_t s);
int assemble_features_c(int nfeat, const int* kind, const int* a, const int* b, const Geom& g,
                        const i64* band_idx, const i64* n_band, const double* planes,
                        i64 plane_stride, const double* item_feat, const double* com, double* X,
                        int origin0, cudaStream_t s);
int band_sums(i64 voxels, i64 batch, const double* planes, i64 plane_stride, int nplanes,
              const i64* band_idx, const i64* item_offsets, int mode, const double* mean,
              double* sums, void* workspace, cudaStream_t s);

// regress.cu (X: materialised (B, F) matrix; *_band: the feature program is the source)
struct BandFeatures;
int fill_program(FeatProgram& p, int nfeat, const int* kind, const int* a, const int* b);
BandFeatures make_band_features(const FeatProgram& prog, const Geom& g, const i64* band_idx,
                                const double* planes, i64 plane_stride, const double* item_feat,
                                const double* com, int origin0);
int linear_predict(const double* X, const i64* n_band, int F, const double* mean,
                   const double* scale, const double* coef, double intercept, double* out,
                   cudaStream_t s);
int linear_predict_band(const BandFeatures& bf, const i64* n_band, const double* mean,
                        const double* scale, const double* coef, double intercept, double* out,
                        cudaStream_t s);
int forest_predict(const double* X, const i64* n_band, int F, const double* mean,
                   const double* scale, const void* nodes, const double* leaves,
                   const unsigned* root_ref, const u8* root_is_leaf, int n_trees, double divisor,
                   const void* nodes8, const int* tree_base, int max_nodes, double* out,
                   cudaStream_t s);
int forest_predict_band(const BandFeatures& bf, const i64* n_band, const double* mean,
                        const double* scale, const void* nodes, const double* leaves,
                        const unsigned* root_ref, const u8* root_is_leaf, int n_trees,
                        double divisor, const void* nodes8, const int* tree_base, int max_nodes,
                        double* out, cudaStream_t s);
// rank-quantised forest (regress.cu: forest_ranked_kernel; lsml_b200.h: lsml_ranked_forest)
struct RankedArgs {
    const unsigned* words;
    const int* tree_start;
    const u8* root_is_leaf;
    const float* thr_tab;
    const int* thr_off;
    int n_trees, nfeat, fbits, rbits, max_tree_words;
    double divisor;
};
int forest_ranked_predict(const double* X, const i64* n_band, int F, const double* mean,
                          const double* scale, const RankedArgs& a, double* out, cudaStream_t s);
int forest_ranked_predict_band(const BandFeatures& bf, const i64* n_band, const double* mean,
                               const double* scale, const RankedArgs& a, double* out,
                               cudaStream_t s);
int mlp_predict(const double* X, const i64* n_band, const double* mean, const double* scale,
                int n_layers, const int* sizes, const double* const* W, const double* const* b,
                int act, double* out, cudaStream_t s);
int mlp_predict_band(const BandFeatures& bf, const i64* n_band, const double* mean,
                     const double* scale, int n_layers, const int* sizes, const double* const* W,
                     const double* const* b, int act, double* out, cudaStream_t s);

// update.cu
int band_update(const Geom& g, double* u, const double* vel, const i64* band_idx,
                const i64* n_band, double step, double* new_u, double* gmag_out, int apply,
                u64* stats, cudaStream_t s);

// fmm.cu
i64 fmm_workspace_bytes(i64 total);
int fmm_workspace_init(i64 total, void* workspace, cudaStream_t s);
int fmm_rebuild(const Geom& g, const double* u, double band, const u64* stats, u8* mask,
                double* dist_out, void* workspace, const i64* prev_band_idx,
                const i64* prev_n_band, cudaStream_t s);
int fmm_slab_begin(const Geom& g, int own_lo, int own_hi, const double* u, u8* mask,
                   void* workspace, const i64* prev_band_idx, const i64* prev_n_band,
                   const int* dense_planes, cudaStream_t s);
int fmm_slab_march(const Geom& g, int own_lo, int own_hi, const double* u, double band, u8* mask,
                   void* workspace, cudaStream_t s);
int fmm_slab_requeue(const Geom& g, int own_lo, int own_hi, double band, const i64* sides_dev,
                     u8* mask, void* workspace, i64* pending_out, cudaStream_t s);
int fmm_slab_finish(const Geom& g, int own_lo, int own_hi, double band, u8* mask, double* dist_out,
                    void* workspace, cudaStream_t s);
int fmm_counters_ptr(i64 total, void* workspace, const i64** out);
i64 fmm_counters_sizes_offset();
i64 fmm_counters_stamps_offset();
i64 fmm_counters_phase_offset();
i64 fmm_counters_failed_offset();
i64 fmm_counters_offset(i64 total);

i64 axis_moments_workspace_bytes(const Geom& g);
int axis_moments(const Geom& g, const double* u, const u64* stats, const double* com, int ncols,
                 const int* cols, const int* axes, const int* orders, int nfeat, double* item_feat,
                 void* workspace, cudaStream_t s);

// mcubes.cu
i64 surface_area_workspace_bytes(const Geom& g);
int surface_area(const Geom& g, const double* u, double* area, void* workspace, cudaStream_t s);

// misc.cu
int ball_init(const Geom& g, double* u, const double* location, double radius,
              const double* centres_dev, const double* radii_dev, cudaStream_t s);
int overlap_counts(i64 voxels, i64 batch, const double* u, const u8* seg, u64* counts,
                   cudaStream_t s);

}  // namespace lsml
