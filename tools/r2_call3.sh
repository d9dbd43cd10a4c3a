#!/bin/bash
# Round 2, GPU call 3 (one GPU): the suite on the current tree, then the A/B of how a large work
# set is dealt to the lanes of the march kernel (FMM_DENSE_FACTOR), rebuilt on the box.
mkdir -p gpurun_out
O=gpurun_out
echo "=== tests"
timeout 600 python -m pytest tests -m gpu -x -q > $O/r2c3_tests.out 2> $O/r2c3_tests.err
echo "exit $?"; tail -n 3 $O/r2c3_tests.out | cut -c1-300
echo "=== variants"
timeout 700 bash tools/r2_fmm_variants.sh "-DFMM_DENSE_FACTOR=1000000000" "-DFMM_DENSE_FACTOR=4" "-DFMM_DENSE_FACTOR=1" > $O/r2c3_variants.out 2>&1
cat $O/r2c3_variants.out | cut -c1-200
