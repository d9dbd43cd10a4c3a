#!/bin/bash
# Round 2, GPU call 12 (one GPU): suite after merging the publish loads, probes, and an ncu capture
# of the banded update kernels at C4 scale (one 2048-crop engine batch) for their real DRAM traffic.
mkdir -p gpurun_out
O=gpurun_out
echo "=== tests"
timeout 700 python -m pytest tests -m gpu -x -q > $O/r2c12_tests.out 2> $O/r2c12_tests.err
echo "exit $?"; tail -n 3 $O/r2c12_tests.out | cut -c1-300
echo "=== fmm"
(timeout 200 python tools/fmm_probe.py 256 12 2>&1 | grep -E "^it  (5|8|11)" | cut -c1-200
 timeout 200 python bench.py --workload c4 --items 2048 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4-2048', d['ms_per_step'], d['phases_ms_per_engine_batch'])") | tee $O/r2c12_fmm.out
echo "=== ncu c4 update"
for k in band_update_kernel band_apply_kernel; do
    timeout 600 ncu --set full --clock-control none -k regex:$k --launch-skip 40 --launch-count 1 \
        -o $O/r2b_c4_$k -f python bench.py --workload c4 --items 2048 --steps 1 --warmup 1 > $O/ncu_r2b_c4_$k.log 2>&1
    echo "ncu $k exit $?"
done
