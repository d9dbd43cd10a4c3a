#!/bin/bash
# Round 2, GPU call 25 (one GPU): CTAs per SM of the march kernel again, on the final structure.
bash tools/r2_fmm_variants.sh "-DFMM_MIN_BLOCKS=3" "-DFMM_MIN_BLOCKS=5" 2>&1 | cut -c1-200
