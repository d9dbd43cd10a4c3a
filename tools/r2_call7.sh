#!/bin/bash
# Round 2, GPU call 7 (one GPU): 1024-thread shapes of the ranked forest kernel, the march without
# hints on a single volume, C2.
mkdir -p gpurun_out
O=gpurun_out
echo "=== forest"
(for v in 12 t2 t1; do
    LSML_FOREST_RK=$v timeout 200 python tools/forest_probe.py 2>&1 | tail -n 1
done) | tee $O/r2c7_forest.out
echo "=== fmm no hints"
(LSML_FMM_NO_HINTS=1 timeout 200 python tools/fmm_probe.py 256 12 2>&1 | grep -E "^it  (5|8|11)|sweep" | cut -c1-400) | tee $O/r2c7_fmm.out
echo "=== c2"
timeout 240 python bench.py --workload c2 --watchdog 200 > $O/r2c7_bench_c2.json 2> $O/r2c7_bench_c2.err
python -c "import json; d=json.loads(open('$O/r2c7_bench_c2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], {k: d[k] for k in d if 'phase' in k})"
timeout 240 python bench.py --workload c2 --features gestalt32 --watchdog 200 > $O/r2c7_bench_c2g.json 2> $O/r2c7_bench_c2g.err
python -c "import json; d=json.loads(open('$O/r2c7_bench_c2g.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], {k: d[k] for k in d if 'phase' in k})"
