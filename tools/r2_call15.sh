#!/bin/bash
# Round 2, GPU call 15 (one GPU): ncu of the march kernel at C4 scale (one 2048-crop engine batch).
mkdir -p gpurun_out
O=gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:fmm_march_kernel --launch-skip 20 --launch-count 1 \
    -o $O/r2b_c4_fmm_march -f python bench.py --workload c4 --items 2048 --steps 1 --warmup 1 > $O/ncu_r2b_c4_fmm_march.log 2>&1
echo "ncu exit $?"; tail -n 3 $O/ncu_r2b_c4_fmm_march.log | cut -c1-300
