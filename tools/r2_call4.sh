#!/bin/bash
# Round 2, GPU call 4 (one GPU): suite with the rank-quantised forest kernel as the default,
# forest kernel variants, the chunked dealing of the march kernel, a short C3 bench.
mkdir -p gpurun_out
O=gpurun_out
echo "=== tests"
timeout 700 python -m pytest tests -m gpu -x -q > $O/r2c4_tests.out 2> $O/r2c4_tests.err
echo "exit $?"; tail -n 25 $O/r2c4_tests.out | cut -c1-300
echo "=== forest"
for v in "" "LSML_FOREST_RK_V=2" "LSML_FOREST_VARIANT=global"; do
    env $v timeout 200 python tools/forest_probe.py 2>&1 | tail -n 1
done | tee $O/r2c4_forest.out
echo "=== fmm (dynamic chunk)"
(timeout 200 python tools/fmm_probe.py 256 12 2>&1 | grep -E "^it  (5|8|11)|sweep" | cut -c1-400
 timeout 200 python bench.py --workload c4 --items 2048 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4-2048', d['ms_per_step'], d['phases_ms_per_engine_batch'])") | tee $O/r2c4_fmm.out
echo "=== c3"
timeout 400 python bench.py --steps 2 --warmup 3 --sections none --no-cpu-baseline > $O/r2c4_bench_c3.json 2> $O/r2c4_bench_c3.err
python -c "import json; d=json.loads(open('$O/r2c4_bench_c3.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['phases_ms_per_step'], d['e2e']['value'])"
tail -n 3 $O/r2c4_bench_c3.err | cut -c1-300
