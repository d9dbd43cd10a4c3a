#!/bin/bash
# Round 2, GPU call 23 (one GPU): suite + smoke + C3 / C4 / C2 after the tail threshold change.
mkdir -p gpurun_out
O=gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > $O/r2c23_tests.out 2> $O/r2c23_tests.err
echo "exit $?"; tail -n 3 $O/r2c23_tests.out | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout 400 python bench.py --steps 3 --warmup 3 --sections none --no-cpu-baseline > $O/r2c23_bench_c3.json 2>/dev/null
python -c "import json; d=json.loads(open('$O/r2c23_bench_c3.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['phases_ms_per_step'], d['e2e']['value'], d['region_a']['value'])"
timeout 200 python bench.py --workload c4 --items 2048 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4-2048', d['ms_per_step'], d['phases_ms_per_engine_batch'], d['rebuilds_not_converged'])"
timeout 240 python bench.py --workload c2 --watchdog 200 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2', d['value'], d['ms_per_step'], d['phases_ms_per_step'])"
