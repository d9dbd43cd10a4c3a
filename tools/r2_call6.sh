#!/bin/bash
# Round 2, GPU call 6 (one GPU): suite on the restructured march kernel (all stencil loads of a
# replay in flight together, publish phases without per-neighbour round trips, no extra fence,
# prefetched sweep header), its A/B switches, ncu capture of the ranked forest kernel, C3 bench.
mkdir -p gpurun_out
O=gpurun_out
echo "=== tests"
timeout 700 python -m pytest tests -m gpu -x -q > $O/r2c6_tests.out 2> $O/r2c6_tests.err
echo "exit $?"; tail -n 5 $O/r2c6_tests.out | cut -c1-300
echo "=== fmm"
(timeout 200 python tools/fmm_probe.py 256 12 2>&1 | grep -E "^it  (5|8|11)|sweep" | cut -c1-400
 timeout 200 python bench.py --workload c4 --items 2048 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4-2048', d['ms_per_step'], d['phases_ms_per_engine_batch'])") | tee $O/r2c6_fmm.out
echo "=== c3"
timeout 400 python bench.py --steps 2 --warmup 3 --sections none --no-cpu-baseline > $O/r2c6_bench_c3.json 2> $O/r2c6_bench_c3.err
python -c "import json; d=json.loads(open('$O/r2c6_bench_c3.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['phases_ms_per_step'], d['e2e']['value'], d['roofline_velocity'])"
tail -n 3 $O/r2c6_bench_c3.err | cut -c1-300
echo "=== forest ncu"
timeout 300 python tools/forest_probe.py 2>&1 | tail -n 1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:forest_ranked_kernel --launch-skip 5 --launch-count 1 \
    -o $O/r2b_forest_ranked -f python tools/forest_probe.py > $O/ncu_r2b_forest_ranked.log 2>&1
echo "ncu exit $?"
echo "=== fmm variants"
timeout 700 bash tools/r2_fmm_variants.sh "-DFMM_EXTRA_FENCE" "-DFMM_PHASE_CLOCKS" 2>&1 | cut -c1-220 | tee $O/r2c6_variants.out
