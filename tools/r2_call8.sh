#!/bin/bash
# Round 2, GPU call 8 (one GPU): why do more warps not speed the ranked forest kernel up?  ncu of the
# 1024-thread x 2 CTA shape and of the two-chains-per-thread shape.
mkdir -p gpurun_out
O=gpurun_out
for v in t2 22; do
    LSML_FOREST_RK=$v timeout 400 ncu --set full --import-source on --clock-control none -k regex:forest_ranked_kernel --launch-skip 5 --launch-count 1 \
        -o $O/r2b_forest_ranked_$v -f python tools/forest_probe.py > $O/ncu_r2b_forest_ranked_$v.log 2>&1
    echo "ncu $v exit $?"
done
