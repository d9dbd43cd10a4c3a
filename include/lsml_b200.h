/*
 * lsml_b200.h -- C ABI of the B200-native lsml level-set hot path (liblsml_b200.so).
 *
 * Plain pointers and sizes only; no torch / C++ types.  Two groups of entry points:
 *
 * (1) DROP-IN symbols: the six functions that lsml/gradient/masked_gradient.py:8-68 binds with
 *     ctypes from `lsml/util/masked_gradient*.so` (built from lsml/util/_cutil/masked_gradient.c).
 *     Same names, same argument order and meaning, HOST pointers, void return.  Copy or symlink
 *     liblsml_b200.so to lsml/util/masked_gradient_b200.so and the reference runs on the GPU
 *     for this path without any source change.
 *
 * (2) DEVICE entry points `lsml_*`: device pointers + a cudaStream_t passed as void*.  They
 *     return 0 on success or a non-zero status; lsml_last_error() gives the message.  Nothing
 *     in this library falls back to the CPU: without a CUDA device every call fails loudly.
 *
 * Grids are C-contiguous (row-major), float64 unless the name says f32; masks are one byte per
 * voxel (C `_Bool` / numpy bool).  `batch` items of identical shape are stored back to back.
 */
#ifndef LSML_B200_H
#define LSML_B200_H

#include <stdbool.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- (1) drop-in for lsml/util/_cutil/masked_gradient.c --------------------------------- */
/* masked_gradient.c:13-16 */
void gradient_centered3d(int m, int n, int p, double *A, bool *mask, double *di, double *dj,
                         double *dk, double *gmag, double deli, double delj, double delk,
                         int normalize);
/* masked_gradient.c:77-80 */
void gradient_centered2d(int m, int n, double *A, bool *mask, double *di, double *dj,
                         double *gmag, double deli, double delj, int normalize);
/* masked_gradient.c:125-128 */
void gradient_centered1d(int m, double *A, bool *mask, double *di, double *gmag, double deli,
                         int normalize);
/* masked_gradient.c:152-154 */
void gmag_os3d(int m, int n, int p, double *A, bool *mask, double *nu, double *gmag,
               double deli, double delj, double delk);
/* masked_gradient.c:228-230 */
void gmag_os2d(int m, int n, double *A, bool *mask, double *nu, double *gmag, double deli,
               double delj);
/* masked_gradient.c:284-286 */
void gmag_os1d(int m, double *A, bool *mask, double *nu, double *gmag, double deli);

/* ---- (2) device entry points ------------------------------------------------------------- */
int lsml_version(void);
const char *lsml_last_error(void);
/* number of CUDA devices visible, or a negative status */
int lsml_device_count(void);
/* kernels launched by this library since load (for bench.py's gpu_launches claim) */
long long lsml_launch_count(void);

/* Dense masked stencils (masked_gradient.c).  shape[ndim], dx[ndim] are HOST arrays.
 * Outputs are written only where mask != 0 (mask == NULL means everywhere). */
int lsml_gmag_os_f64(int ndim, const int *shape, long long batch, const double *A,
                     const uint8_t *mask, const double *nu, double *gmag, const double *dx,
                     void *stream);
int lsml_gmag_os_f32(int ndim, const int *shape, long long batch, const float *A,
                     const uint8_t *mask, const float *nu, float *gmag, const double *dx,
                     void *stream);
/* grads: ndim output planes stored back to back (plane a at grads + a*batch*voxels). */
int lsml_gradient_centered_f64(int ndim, const int *shape, long long batch, const double *A,
                               const uint8_t *mask, double *grads, double *gmag,
                               const double *dx, int normalize, void *stream);
int lsml_gradient_centered_f32(int ndim, const int *shape, long long batch, const float *A,
                               const uint8_t *mask, float *grads, float *gmag,
                               const double *dx, int normalize, void *stream);

/* Band compaction (numpy boolean indexing `X[mask]`, lsml/core/model.py:388,394):
 * writes the ascending row-major flat indices of the non-zero mask bytes of the whole
 * batch (index = item*voxels + local) to band_idx (capacity total) and the count to *n_band
 * (device int64).  item_offsets (device int64[batch+1], may be NULL) receives the CSR offsets
 * of the items inside band_idx.
 * workspace: lsml_compact_workspace_bytes(total) bytes of device memory. */
long long lsml_compact_workspace_bytes(long long total);
int lsml_compact_mask(long long total, long long voxels_per_item, const uint8_t *mask,
                      long long *band_idx, long long *n_band, long long *item_offsets,
                      void *workspace, void *stream);


/* ---- image planes: one-time, u-independent (lsml/feature/provided/image.py) ---------------- */
/* scipy.ndimage.gaussian_filter1d along `axis` (0..ndim-1), mode='reflect', evaluated in
 * scipy's own summation order (bit-identical planes).  weights: DEVICE array of 2*radius+1
 * doubles, already reversed as gaussian_filter1d does (image.py:28-29, 52-56); antisym=1 for
 * order-1 kernels.  mode: 0 out=t, 1 out=t*t, 2 out+=t*t, 3 out=sqrt(out+t*t), 4 out=sqrt(t*t)
 * (modes 1-4 build ImageEdgeSample's magnitude without temporaries, image.py:59-61). */
int lsml_gauss1d_f64(int ndim, const int *shape, long long batch, const double *in, double *out,
                     const double *weights, int radius, int antisym, int axis, int mode,
                     void *stream);
/* scipy.ndimage.gaussian_filter(in, (sigma0, sigma1)) on a batch of 2-D items small enough to
 * live in shared memory twice (both passes in one kernel, bit-identical to two lsml_gauss1d_f64
 * calls).  w0 / w1: DEVICE weights with the centre at w[r]; r < 0 skips that axis.  *done = 0
 * when the items are too large: chain lsml_gauss1d_f64 instead.  (examples/gestalt2d/
 * prev_iter_feature.py:21-22 calls gaussian_filter(u, sigma) every iteration.) */
int lsml_gauss2d_f64(const int *shape, long long batch, const double *in, double *out,
                     const double *w0, int r0, const double *w1, int r1, int *done, void *stream);
/* |numpy.gradient(img, *dx)| (ImageEdgeSample with sigma == 0, image.py:48-49) */
int lsml_np_gradient_mag_f64(int ndim, const int *shape, long long batch, const double *in,
                             double *out, const double *dx, void *stream);
/* (img - img.mean()) / img.std() per item (lsml/core/model.py:328-329) */
long long lsml_normalize_workspace_bytes(long long voxels, long long batch);
int lsml_normalize_f64(long long voxels, long long batch, const double *img, double *out,
                       void *workspace, void *stream);

/* ---- per-iteration feature extraction ------------------------------------------------------- */
/* Exact integer statistics of {u > 0} per item (Size / Moments / centre of mass, shape.py:28,
 * 201-230): stats[item*8 + {0: count, 1..3: sum of index per padded axis, 4..6: sum of index^2,
 * 7: number of exact zeros}] as uint64. */
int lsml_global_stats(int ndim, const int *shape, long long batch, const double *u,
                      unsigned long long *stats, void *stream);
/* mean and population std of `nplanes` image planes over each item's band
 * (InteriorImageAverage / InteriorImageVariation, image.py:78-121): out arrays [batch][nplanes] */
long long lsml_band_stats_workspace_bytes(long long voxels, long long batch, int nplanes);
int lsml_band_stats(long long voxels, long long batch, const double *planes,
                    long long plane_stride, int nplanes, const long long *band_idx,
                    const long long *item_offsets, double *mean, double *stdv, void *workspace,
                    void *stream);
/* BoundarySize for 2-D items: marching-squares length of {u = 0} (shape.py:70-85) */
long long lsml_contour_workspace_bytes(int m, int n, long long batch);
int lsml_contour_length(int ndim, const int *shape, long long batch, const double *u,
                        const double *dx, double *boundary, void *workspace, void *stream);
/* Feature program: column f has kind[f] in {0 plane sample (a=plane), 1 distance to centre of
 * mass, 2 band mean (a=plane), 3 band std (a=plane), 4 size, 5 boundary size, 6 isoperimetric
 * ratio, 7 moment (a=axis, b=order), 8 COMRaySample column (image.py:124-171; a=plane holding
 * the raw image, b = sample index j | n_samples << 16; 2-D/3-D, not on slabs)}.  kind/a/b are
 * HOST int arrays of length nfeat (<= 64). */
int lsml_item_features(int nfeat, const int *kind, const int *a, const int *b, int ndim,
                       const int *shape, long long batch, const double *dx, int nplanes,
                       const unsigned long long *stats, const double *band_mean,
                       const double *band_std, const double *boundary, double *item_feat,
                       double *com, void *stream);
/* FeatureMap.__call__(...)[mask] (feature_map.py:53-80): X is (n_band, nfeat) row-major */
int lsml_assemble_features(int nfeat, const int *kind, const int *a, const int *b, int ndim,
                           const int *shape, long long batch, const double *dx,
                           const long long *band_idx, const long long *n_band,
                           const double *planes, long long plane_stride,
                           const double *item_feat, const double *com, double *X, void *stream);

/* The dense float64 3-D gmag_os runs as a TMA plane pipeline (cp.async.bulk.tensor + mbarrier
 * stages, csrc/stencil.cu) when the grid is eligible; on = 0 forces the register-marching
 * generation instead, on = 1 re-enables it, on < 0 only queries.  Returns the previous setting.
 * Both generations are bit-exact; the switch exists for A/B measurements and tests. */
int lsml_stencil_use_tma(int on);

/* Moments of order > 2 (shape.py:214-230 takes any order >= 1): from exact per-axis histograms of
 * {u > 0}; writes item_feat[item*nfeat + cols[q]] for the ncols requested (axis, order) pairs.
 * Call after lsml_item_features (needs com); cols/axes/orders are DEVICE int arrays. */
long long lsml_axis_moments_workspace_bytes(int ndim, const int *shape, long long batch);
int lsml_axis_moments(int ndim, const int *shape, long long batch, const double *dx,
                      const double *u, const unsigned long long *stats, const double *com,
                      int ncols, const int *cols, const int *axes, const int *orders, int nfeat,
                      double *item_feat, void *workspace, void *stream);

/* BoundarySize, ndim = 3 (shape.py:88-89: skimage marching_cubes_lewiner + mesh_surface_area):
 * area[item] = surface area of {u = 0} in physical units, a marching-cubes reduction without a
 * mesh (csrc/mcubes.cu; third-party algorithm, parity unpinned beyond test_shape.py:59-107). */
long long lsml_surface_area_workspace_bytes(int ndim, const int *shape, long long batch);
int lsml_surface_area(int ndim, const int *shape, long long batch, const double *u,
                      const double *dx, double *area, void *workspace, void *stream);

/* ---- regressor inference (model.py:388; parameters exported from scikit-learn) -------------- */
/* mean/scale: StandardScaler parameters or NULL. */
int lsml_linear_predict(const double *X, const long long *n_band, int nfeat, const double *mean,
                        const double *scale, const double *coef, double intercept, double *out,
                        void *stream);
/* nodes: 32-byte super nodes (an internal node and its two children per entry, see regress.cu;
 * 32-byte aligned); leaves: float64 leaf values; root_ref: unsigned[n_trees], the root super
 * node of each tree or -- root_is_leaf[t] != 0 -- the leaf index of a single-leaf tree;
 * divisor = n_estimators.  nodes8 / tree_base / max_nodes (optional, NULL / 0): the same forest
 * as 8-byte single nodes, int[n_trees + 1] first node of each tree then the node count, and the
 * size of the largest tree -- forests of small trees (max_nodes <= 2048) are then walked from
 * shared memory, tree by tree (regress.cu: forest_staged_kernel). */
int lsml_forest_predict(const double *X, const long long *n_band, int nfeat, const double *mean,
                        const double *scale, const void *nodes, const double *leaves,
                        const unsigned *root_ref, const uint8_t *root_is_leaf, int n_trees,
                        double divisor, const void *nodes8, const int *tree_base, int max_nodes,
                        double *out, void *stream);
/* sizes: HOST int[n_layers+1]; W, b: HOST arrays of n_layers DEVICE pointers;
 * act: 0 identity, 1 relu, 2 tanh, 3 logistic */
int lsml_mlp_predict(const double *X, const long long *n_band, const double *mean,
                     const double *scale, int n_layers, const int *sizes,
                     const double *const *W, const double *const *b, int act, double *out,
                     void *stream);

/* The same three regressors with the FEATURE PROGRAM as their input instead of a materialised
 * matrix: row b of the virtual matrix is FeatureMap(...)[band voxel b] (feature_map.py:53-80),
 * gathered from the image planes and per-item scalars inside the kernel and never written to
 * memory.  This is what the evolution loop (model.py:384-388) runs.  kind/a/b, shape, dx: HOST
 * arrays as for lsml_assemble_features; the other pointers are DEVICE arrays. */
typedef struct lsml_band_source {
    int nfeat;
    const int *kind, *a, *b;
    int ndim;
    const int *shape;
    long long batch;
    const double *dx;
    const long long *band_idx;
    const double *planes;
    long long plane_stride;
    const double *item_feat;
    const double *com;
    int origin0;          /* first axis-0 plane of a slab's local grid in the whole volume, else 0 */
} lsml_band_source;
int lsml_linear_predict_band(const lsml_band_source *src, const long long *n_band,
                             const double *mean, const double *scale, const double *coef,
                             double intercept, double *out, void *stream);
int lsml_forest_predict_band(const lsml_band_source *src, const long long *n_band,
                             const double *mean, const double *scale, const void *nodes,
                             const double *leaves, const unsigned *root_ref,
                             const uint8_t *root_is_leaf, int n_trees, double divisor,
                             const void *nodes8, const int *tree_base, int max_nodes, double *out,
                             void *stream);
int lsml_mlp_predict_band(const lsml_band_source *src, const long long *n_band, const double *mean,
                          const double *scale, int n_layers, const int *sizes,
                          const double *const *W, const double *const *b, int act, double *out,
                          void *stream);

/* A forest in the RANK-QUANTISED layout (lsml_b200/core/regressors.py: pack_ranked; regress.cu:
 * forest_ranked_kernel): per feature f the sorted distinct float32 thresholds S_f of the forest
 * (thr_tab = S_0 | S_1 | ..., thr_off = int[nfeat + 1]); trees in preorder, an internal node one
 * 32-bit word, fields from the top  feature (fbits) | j (rbits) | offset of the right child in
 * words | bit 1 / bit 0: the right / left child is a leaf  with threshold S_f[j], the left
 * child in the next word, a leaf as its float64 value in two words; tree_start = int[n_trees + 1]
 * word offsets (multiples of 4; words 16-byte aligned); max_tree_words = the largest tree.  Same
 * decisions, same leaves, same float64 sum in tree order as lsml_forest_predict -- i.e. as
 * scikit-learn's predict (model.py:388) -- walked from shared memory.  All pointers DEVICE. */
typedef struct lsml_ranked_forest {
    const unsigned *words;
    const int *tree_start;
    const uint8_t *root_is_leaf;
    const float *thr_tab;
    const int *thr_off;
    int n_trees, nfeat, fbits, rbits, max_tree_words;
    double divisor;
} lsml_ranked_forest;
int lsml_forest_ranked_predict(const double *X, const long long *n_band, int nfeat,
                               const double *mean, const double *scale,
                               const lsml_ranked_forest *forest, double *out, void *stream);
int lsml_forest_ranked_predict_band(const lsml_band_source *src, const long long *n_band,
                                    const double *mean, const double *scale,
                                    const lsml_ranked_forest *forest, double *out, void *stream);

/* ---- banded update u[mask] += step*v*|Du|_upwind (model.py:390-394) ------------------------- */
/* new_u[b] receives the updated value of band voxel b; with apply != 0 it is also scattered
 * into u.  gmag_out (may be NULL) receives the upwind gradient magnitude per band voxel. */
int lsml_band_update(int ndim, const int *shape, long long batch, const double *dx, double *u,
                     const double *vel, const long long *band_idx, const long long *n_band,
                     double step, double *new_u, double *gmag_out, int apply, void *stream);

/* Same with apply, and the lsml_global_stats record of u (stats, device) brought up to date from
 * the voxels whose sign class changed -- exactly what a dense recount of the new u would give. */
int lsml_band_update_stats(int ndim, const int *shape, long long batch, const double *dx, double *u,
                           const double *vel, const long long *band_idx, const long long *n_band,
                           double step, double *new_u, double *gmag_out,
                           unsigned long long *stats, void *stream);

/* ---- band rebuild: distance_transform(u, band, dx) (lsml/util/distance_transform.py:5-59) --- */
long long lsml_fmm_workspace_bytes(long long total);
int lsml_fmm_workspace_init(long long total, void *workspace, void *stream);
/* mask (uint8, zero on first use) and workspace persist between calls; stats = lsml_global_stats
 * of the same u; dist_out may be NULL (0 outside the band otherwise). */
int lsml_fmm_rebuild(int ndim, const int *shape, long long batch, const double *dx,
                     const double *u, double band, const unsigned long long *stats,
                     uint8_t *mask, double *dist_out, void *workspace, void *stream);
/* Same, for the evolution loop (model.py:394-397): u has changed ONLY on the band of the
 * previous rebuild (prev_band_idx[0 .. *prev_n_band), device arrays as written by
 * lsml_compact_mask), so the interface is located from that band instead of a dense pass. */
int lsml_fmm_rebuild_from_band(int ndim, const int *shape, long long batch, const double *dx,
                               const double *u, double band, const unsigned long long *stats,
                               uint8_t *mask, double *dist_out, void *workspace,
                               const long long *prev_band_idx, const long long *prev_n_band,
                               void *stream);
/* ---- slab decomposition of ONE large volume along axis 0 (host side:
 * lsml_b200/core/slab.py).  The local grid `shape` is the rank's owned planes [own_lo, own_hi)
 * plus halo planes that hold copies of the neighbour ranks' voxels.  A rebuild is
 *   begin -> { march -> exchange the halo planes of T (first 8*total bytes of the workspace) and
 *   of mask -> requeue } until no rank has pending work -> finish.
 * The fixed point is the same as the single-GPU one, so masks and distances are identical. */
/* [own_lo, own_hi) here is the range of planes the rank EVALUATES (its owned planes plus the
 * ghost planes it solves redundantly); dense_planes = HOST int[4] {a0, b0, a1, b1}: plane ranges
 * searched densely for the interface besides the previous band (prev_band_idx != NULL only). */
int lsml_fmm_slab_begin(int ndim, const int *shape, const double *dx, int own_lo, int own_hi,
                        const double *u, uint8_t *mask, void *workspace,
                        const long long *prev_band_idx, const long long *prev_n_band,
                        const int *dense_planes, void *stream);
int lsml_fmm_slab_march(int ndim, const int *shape, const double *dx, int own_lo, int own_hi,
                        const double *u, double band, uint8_t *mask, void *workspace,
                        void *stream);
/* sides_dev: DEVICE int64[2], non-zero = the halo below / above the owned planes changed in the
 * exchange (kept on the device so that a round needs a single host read);
 * pending_out: DEVICE int64, number of voxels queued for the next march (0 = this rank is done) */
int lsml_fmm_slab_requeue(int ndim, const int *shape, const double *dx, int own_lo, int own_hi,
                          double band, const long long *sides_dev, uint8_t *mask,
                          void *workspace, long long *pending_out, void *stream);
int lsml_fmm_slab_finish(int ndim, const int *shape, const double *dx, int own_lo, int own_hi,
                         double band, uint8_t *mask, double *dist_out, void *workspace,
                         void *stream);
/* compaction of a sub-range of a grid: indices are written with index_offset added */
int lsml_compact_mask_range(long long total, const uint8_t *mask, long long *band_idx,
                            long long *n_band, long long index_offset, void *workspace,
                            void *stream);
/* raw band sums per item and plane: mode 0 sum x, mode 1 sum (x - mean)^2 (mean: [batch][nplanes]) */
int lsml_band_sums(long long voxels, long long batch, const double *planes,
                   long long plane_stride, int nplanes, const long long *band_idx,
                   const long long *item_offsets, int mode, const double *mean, double *sums,
                   void *workspace, void *stream);
/* lsml_assemble_features for a local grid whose first axis-0 plane is plane origin0 of the volume */
int lsml_assemble_features_origin(int nfeat, const int *kind, const int *a, const int *b, int ndim,
                                  const int *shape, long long batch, const double *dx,
                                  const long long *band_idx, const long long *n_band,
                                  const double *planes, long long plane_stride,
                                  const double *item_feat, const double *com, double *X,
                                  int origin0, void *stream);

/* synchronous: copies the 8 int64 counters of the last rebuild to HOST out
 * {n_init, n_cand, single-CTA sweeps, sweeps, converged, prev_init, prev_cand, replays} */
int lsml_fmm_counters(long long total, void *workspace, long long *out, void *stream);
/* diagnostics: work-set size of the first 32 grid-wide sweeps of the last rebuild, followed by
 * their start times in ns (HOST out[64]) */
int lsml_fmm_sweep_sizes(long long total, void *workspace, long long *out, void *stream);

/* Status words a host layer can read without a call: the workspace holds the rebuild counters at
 * byte offset lsml_fmm_counters_offset(total); at lsml_fmm_status_offset() bytes into them lie
 * int64 {failed_rebuilds (rebuilds since workspace_init that hit the sweep limit before their
 * fixed point -- must stay 0), barriers of the last march, live sweep, live rix, live begin,
 * live end (published before every grid-wide sweep; readable from a second stream while a march
 * runs)}. */
long long lsml_fmm_counters_offset(long long total);
long long lsml_fmm_status_offset(void);
/* diagnostics (builds with -DFMM_PHASE_CLOCKS): byte offset, inside the counters, of int64[8] SM
 * cycles one lane spent per phase of the small grid-wide sweeps (see FmmCounters::phase) */
long long lsml_fmm_phase_offset(void);

/* ---- either side of the path -------------------------------------------------- */
/* u0 = 2*ball - 1 with ball = (radius - |idx*dx - location*dx|) > 0
 * (lsml/initializer/provided/ball.py:13-31, initializer_base.py:56).  location: HOST double[ndim]
 * in index units (used when centres == NULL); centres / radii: optional DEVICE per-item arrays
 * (centres[item*3 + padded axis], physical units; radii[item]). */
int lsml_ball_init(int ndim, const int *shape, long long batch, const double *dx, double *u,
                   const double *location, double radius, const double *centres,
                   const double *radii, void *stream);
/* counts[item*2 + {0,1}] = |{u>0} & seg|, |{u>0} | seg|  (lsml/score_functions.py:4-30) */
int lsml_overlap_counts(long long voxels, long long batch, const double *u, const uint8_t *seg,
                        unsigned long long *counts, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LSML_B200_H */
