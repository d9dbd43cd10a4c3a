#!/bin/bash
# Round 2, GPU call 5 (one GPU): shapes of the ranked forest kernel (voxels per thread x CTAs per
# SM), walk-depth statistics, hint slack in the march scheduler.
mkdir -p gpurun_out
O=gpurun_out
echo "=== forest tests"
timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -x -q -k "regressors or ranked or full_step" 2>&1 | tail -n 3
echo "=== forest"
(LSML_FOREST_DEPTHS=1 timeout 300 python tools/forest_probe.py 2>&1 | tail -n 2
for v in 13 22 23 21; do
    LSML_FOREST_RK=$v timeout 200 python tools/forest_probe.py 2>&1 | tail -n 1
done) | tee $O/r2c5_forest.out
echo "=== fmm hint slack"
timeout 900 bash tools/r2_fmm_variants.sh "-DFMM_PHASE_CLOCKS" "-DFMM_CHUNK_CEIL" "-DFMM_HINT_SLACK_SHIFT=3" "-DFMM_HINT_SLACK_SHIFT=2" 2>&1 | cut -c1-200 | tee $O/r2c5_fmm.out
