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
 * (device int64).  item_counts (device int64[batch], may be NULL) receives per-item counts.
 * workspace: lsml_compact_workspace_bytes(total) bytes of device memory. */
long long lsml_compact_workspace_bytes(long long total);
int lsml_compact_mask(long long total, long long voxels_per_item, const uint8_t *mask,
                      long long *band_idx, long long *n_band, long long *item_counts,
                      void *workspace, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LSML_B200_H */
