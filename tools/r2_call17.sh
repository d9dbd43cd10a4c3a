#!/bin/bash
# Round 2, GPU call 17 (one GPU): interior fast paths of replay / publish (no range tests for voxels two
# cells inside every face): suite, probes, C3 bench.
mkdir -p gpurun_out
O=gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > $O/r2c17_tests.out 2> $O/r2c17_tests.err
echo "exit $?"; tail -n 3 $O/r2c17_tests.out | cut -c1-300
(timeout 200 python tools/fmm_probe.py 256 12 2>&1 | grep -E "^it  (5|8|11)|sweep us" | cut -c1-400
 timeout 200 python bench.py --workload c4 --items 2048 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4-2048', d['ms_per_step'], d['phases_ms_per_engine_batch'])")
timeout 400 python bench.py --steps 2 --warmup 3 --sections none --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['phases_ms_per_step'])"
