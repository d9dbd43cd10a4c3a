#!/bin/bash
# Round 2, GPU call 20 (one GPU): hint slack on batches (C4, C2) against LSML_FMM_NO_SLACK=1, the
# suite (incl. the new marched-axes plane test).
mkdir -p gpurun_out
O=gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > $O/r2c20_tests.out 2> $O/r2c20_tests.err
echo "exit $?"; tail -n 3 $O/r2c20_tests.out | cut -c1-300
for e in "" "LSML_FMM_NO_SLACK=1"; do
  echo "== slack ${e:-on}"
  env $e timeout 200 python bench.py --workload c4 --items 2048 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4-2048', d['ms_per_step'], d['phases_ms_per_engine_batch']['rebuild'], d['rebuilds_not_converged'])"
  env $e timeout 240 python bench.py --workload c2 --watchdog 200 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2', d['value'], d['ms_per_step'], d['phases_ms_per_step']['rebuild'])"
  env $e timeout 240 python bench.py --workload c2 --features gestalt32 --watchdog 200 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2g', d['value'], d['ms_per_step'], d['phases_ms_per_step']['rebuild'])"
done
