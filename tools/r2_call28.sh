#!/bin/bash
# Round 2, GPU call 28 (one GPU): fewer, larger CTAs for the march kernel (same warps per SM, fewer barrier
# participants).
bash tools/r2_fmm_variants.sh "-DFMM_THREADS_PER_CTA=384 -DFMM_MIN_BLOCKS=1" "-DFMM_THREADS_PER_CTA=192 -DFMM_MIN_BLOCKS=2" 2>&1 | cut -c1-200
