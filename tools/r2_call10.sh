#!/bin/bash
# Round 2, GPU call 10 (one GPU): bulk-copied tree groups in the ranked forest kernel (suite, probe
# shapes), short C3 / C4 / C2 benches.
mkdir -p gpurun_out
O=gpurun_out
echo "=== tests"
timeout 700 python -m pytest tests -m gpu -x -q > $O/r2c10_tests.out 2> $O/r2c10_tests.err
echo "exit $?"; tail -n 4 $O/r2c10_tests.out | cut -c1-300
echo "=== forest"
(for v in "" 12 t2; do
    LSML_FOREST_RK=$v timeout 200 python tools/forest_probe.py 2>&1 | tail -n 1
done) | tee $O/r2c10_forest.out
echo "=== c3"
timeout 400 python bench.py --steps 2 --warmup 3 --sections none --no-cpu-baseline > $O/r2c10_bench_c3.json 2> $O/r2c10_bench_c3.err
python -c "import json; d=json.loads(open('$O/r2c10_bench_c3.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['phases_ms_per_step'], d['e2e']['value'], d['region_a'])"
echo "=== c4 / c2"
timeout 200 python bench.py --workload c4 --items 2048 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4-2048', d['ms_per_step'], d['phases_ms_per_engine_batch'])"
timeout 240 python bench.py --workload c2 --watchdog 200 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2', d['value'], d['ms_per_step'], {k: d[k] for k in d if 'phase' in k})"
