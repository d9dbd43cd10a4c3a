#!/bin/bash
# Round 2, GPU call 27 (one GPU): the bench line of the final build (N = 1, with its sections).
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python bench.py --steps 3 --warmup 3 --verbose --watchdog 1000 > $O/r2b_bench_n1.json 2> $O/r2b_bench_n1.err
echo "bench exit $?"
python -c "import json; d=json.loads(open('$O/r2b_bench_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['phases_ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['region_a']['value'], d['c4']['value'], d['c4']['ms_per_step'], d['c4']['phases_ms_per_engine_batch'], d['c5']['value'], d['c5']['ms_per_iteration'], d['cpu_baseline']['value'], d['roofline']['ms'])"
timeout 200 python tools/fmm_probe.py 256 12 > $O/r2b_fmm_probe.log 2>&1; grep -E "^it  (5|8)" $O/r2b_fmm_probe.log | cut -c1-160
