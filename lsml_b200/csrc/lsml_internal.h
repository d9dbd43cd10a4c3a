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

// compact.cu
i64 compact_workspace_bytes(i64 total);
int compact_mask(i64 total, const u8* mask, i64* band_idx, i64* n_band, void* workspace,
                 cudaStream_t s);
int item_offsets(const i64* band_idx, const i64* n_band, i64 voxels, i64 batch, i64* offsets,
                 cudaStream_t s);

}  // namespace lsml
