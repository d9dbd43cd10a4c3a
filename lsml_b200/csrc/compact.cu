// compact.cu -- narrow-band compaction.
//
// The reference selects band voxels with numpy boolean indexing (`features[mask]`,
// `velocity[mask] = ...`, lsml/core/model.py:388,394), i.e. in ascending row-major flat-index
// order; that order is the row order of the regression matrix and of `predict`'s output, so it
// must be reproduced exactly.  Integer work, bit-exact by construction.
//
// Three launches (reduce - scan - scatter): a tile is 256 threads x 16 mask bytes (one 128-bit
// load per thread, 4 KiB per CTA).  Pass 1 counts the non-zero bytes per tile (popc of a
// byte-wise != 0 bit mask, warp shuffle + block sum), pass 2 is an exclusive scan of the tile
// counts by one CTA, pass 3 re-reads the tile (L2-resident for grids up to ~100 MiB of mask),
// scans the 256 per-thread counts with warp shuffles and writes the indices in order.
// Algorithmic bytes: 1*G (mask) + 8*B (indices); traffic is 2*G + 8*B when the mask does not
// fit L2.
#include "common.cuh"

namespace lsml {

constexpr int CT_THREADS = 256;
constexpr int CT_BYTES_PER_THREAD = 16;
constexpr int CT_TILE = CT_THREADS * CT_BYTES_PER_THREAD;

// 16-bit set: bit b is set when byte b of this thread's 16-byte chunk is non-zero
__device__ __forceinline__ unsigned chunk_bits(const u8* __restrict__ mask, i64 start, i64 total,
                                               bool aligned) {
    unsigned bits = 0;
    if (start >= total) return 0;
    if (aligned && start + CT_BYTES_PER_THREAD <= total) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(mask + start));
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int b = 0; b < 4; ++b)
                bits |= ((w[q] >> (8 * b)) & 0xffu) ? (1u << (4 * q + b)) : 0u;
        }
    } else {
        for (int b = 0; b < CT_BYTES_PER_THREAD && start + b < total; ++b)
            bits |= mask[start + b] ? (1u << b) : 0u;
    }
    return bits;
}

__global__ void __launch_bounds__(CT_THREADS)
compact_count_kernel(const u8* __restrict__ mask, i64 total, i64* __restrict__ tile_counts,
                     bool aligned) {
    __shared__ int red[32];
    const i64 start = ((i64)blockIdx.x * CT_THREADS + threadIdx.x) * CT_BYTES_PER_THREAD;
    const int c = __popc(chunk_bits(mask, start, total, aligned));
    const int s = block_sum<int>(c, red);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = s;
}

// exclusive scan of tile_counts[0..n) in place by one CTA of 1024 threads; total -> *n_band
__global__ void __launch_bounds__(1024)
compact_scan_kernel(i64* __restrict__ tile_counts, i64 n, i64* __restrict__ n_band) {
    __shared__ i64 warp_tot[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const i64 per = (n + 1023) / 1024;
    const i64 lo = min(n, (i64)tid * per), hi = min(n, lo + per);
    i64 sum = 0;
    for (i64 i = lo; i < hi; ++i) sum += tile_counts[i];
    // inclusive warp scan of the per-thread sums
    i64 inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const i64 t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        i64 w = warp_tot[lane];
        i64 winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const i64 t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        warp_tot[lane] = winc - w;   // exclusive warp offsets
        if (lane == 31) *n_band = winc;
    }
    __syncthreads();
    i64 run = warp_tot[warp] + (inc - sum);
    for (i64 i = lo; i < hi; ++i) {
        const i64 c = tile_counts[i];
        tile_counts[i] = run;
        run += c;
    }
}

__global__ void __launch_bounds__(CT_THREADS)
compact_scatter_kernel(const u8* __restrict__ mask, i64 total,
                       const i64* __restrict__ tile_offsets, i64* __restrict__ band_idx,
                       bool aligned, i64 index_offset) {
    __shared__ int warp_tot[CT_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const i64 start = ((i64)blockIdx.x * CT_THREADS + tid) * CT_BYTES_PER_THREAD;
    unsigned bits = chunk_bits(mask, start, total, aligned);
    const int c = __popc(bits);
    int inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    int woff = 0;
#pragma unroll
    for (int w = 0; w < CT_THREADS / 32; ++w) woff += (w < warp) ? warp_tot[w] : 0;
    i64 pos = tile_offsets[blockIdx.x] + woff + (inc - c);
    while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1;
        band_idx[pos++] = start + b + index_offset;
    }
}

// item_offsets[i] = first position in band_idx whose value >= i*voxels (i = 0..batch)
__global__ void item_offsets_kernel(const i64* __restrict__ band_idx,
                                    const i64* __restrict__ n_band, i64 voxels, i64 batch,
                                    i64* __restrict__ item_offsets) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > batch) return;
    const i64 n = *n_band, key = i * voxels;
    i64 lo = 0, hi = n;
    while (lo < hi) {
        const i64 mid = (lo + hi) >> 1;
        if (band_idx[mid] < key) lo = mid + 1; else hi = mid;
    }
    item_offsets[i] = lo;
}

i64 compact_workspace_bytes(i64 total) {
    const i64 tiles = (total + CT_TILE - 1) / CT_TILE;
    return (tiles + 1) * (i64)sizeof(i64);
}

// index_offset is added to every index written (the mask may be a sub-range of a larger grid:
// the owned planes of a slab)
int compact_mask(i64 total, const u8* mask, i64* band_idx, i64* n_band, void* workspace,
                 i64 index_offset, cudaStream_t s) {
    if (total <= 0) {
        LSML_CUDA(cudaMemsetAsync(n_band, 0, sizeof(i64), s));
        return 0;
    }
    const i64 tiles = (total + CT_TILE - 1) / CT_TILE;
    i64* tile_counts = static_cast<i64*>(workspace);
    const bool aligned = (reinterpret_cast<uintptr_t>(mask) & 15) == 0;
    compact_count_kernel<<<(unsigned)tiles, CT_THREADS, 0, s>>>(mask, total, tile_counts, aligned);
    LSML_LAUNCH_CHECK("compact_count_kernel");
    compact_scan_kernel<<<1, 1024, 0, s>>>(tile_counts, tiles, n_band);
    LSML_LAUNCH_CHECK("compact_scan_kernel");
    compact_scatter_kernel<<<(unsigned)tiles, CT_THREADS, 0, s>>>(mask, total, tile_counts, band_idx, aligned,
                                                                  index_offset);
    LSML_LAUNCH_CHECK("compact_scatter_kernel");
    return 0;
}

int item_offsets(const i64* band_idx, const i64* n_band, i64 voxels, i64 batch, i64* offsets,
                 cudaStream_t s) {
    const int threads = 256;
    item_offsets_kernel<<<(unsigned)((batch + 1 + threads - 1) / threads), threads, 0, s>>>(
        band_idx, n_band, voxels, batch, offsets);
    LSML_LAUNCH_CHECK("item_offsets_kernel");
    return 0;
}

}  // namespace lsml
