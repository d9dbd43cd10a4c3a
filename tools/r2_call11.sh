#!/bin/bash
# Round 2, GPU call 11 (one GPU): smoke, the default bench line (N = 1 with its c4 / c5 sections
# and the CPU baseline leg), then the profiling pass.
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout 1200 python bench.py --steps 3 --warmup 3 --verbose --watchdog 1000 > $O/r2b_bench_n1.json 2> $O/r2b_bench_n1.err
echo "bench exit $?"; tail -n 6 $O/r2b_bench_n1.err | cut -c1-300
python -c "import json; d=json.loads(open('$O/r2b_bench_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['phases_ms_per_step'], d['e2e']['value'], d['region_a']['value'], d['c4']['value'], d['c4']['ms_per_step'], d['c5']['value'], d['c5']['ms_per_iteration'], d['cpu_baseline']['value'])"
timeout 900 bash tools/r2b_profiles.sh > $O/r2b_profiles.log 2>&1
echo "profiles exit $?"; tail -n 12 $O/r2b_profiles.log | cut -c1-300
