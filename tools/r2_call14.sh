#!/bin/bash
# Round 2, GPU call 14 (one GPU): per-item feature columns ranked once per warp in the forest kernel.
mkdir -p gpurun_out
O=gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > $O/r2c14_tests.out 2> $O/r2c14_tests.err
echo "exit $?"; tail -n 3 $O/r2c14_tests.out | cut -c1-300
timeout 200 python tools/forest_probe.py 2>&1 | tail -n 1
timeout 400 python bench.py --steps 2 --warmup 3 --sections none --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['phases_ms_per_step'], d['region_a']['value'])"
timeout 200 python bench.py --workload c4 --items 2048 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4-2048', d['ms_per_step'], d['phases_ms_per_engine_batch'])"
timeout 240 python bench.py --workload c2 --watchdog 200 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2', d['value'], d['ms_per_step'], {k: d[k] for k in d if 'phase' in k})"
