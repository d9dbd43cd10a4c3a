#!/bin/bash
# Round 2, GPU call 9 (one GPU): the pipe-balanced ranked forest walk (new node layout), early hints.
mkdir -p gpurun_out
O=gpurun_out
echo "=== forest tests"
timeout 300 python -m pytest tests/test_engine_gpu.py tests/test_baseline_sizes_gpu.py -m gpu -x -q 2>&1 | tail -n 3
echo "=== forest"
(for v in 12 t2 t1 22; do
    LSML_FOREST_RK=$v timeout 200 python tools/forest_probe.py 2>&1 | tail -n 1
done) | tee $O/r2c9_forest.out
echo "=== fmm early hints"
timeout 900 bash tools/r2_fmm_variants.sh "-DFMM_HINT_EARLY=1" "-DFMM_HINT_EARLY=2" "-DFMM_HINT_EARLY=3" 2>&1 | cut -c1-220 | tee $O/r2c9_variants.out
