#!/bin/bash
# Round 2, GPU call 16 (one GPU): code-size variants of the march kernel (cold fallback as a call;
# every call site as a call).
mkdir -p gpurun_out
O=gpurun_out
echo "== default (cold fallback through the call)"
(timeout 200 python tools/fmm_probe.py 256 12 2>&1 | grep -E "^it  (5|8|11)" | cut -c1-200
 timeout 200 python bench.py --workload c4 --items 2048 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4-2048', d['ms_per_step'], d['phases_ms_per_engine_batch']['rebuild'])")
timeout 700 bash tools/r2_fmm_variants.sh "-DFMM_CALL_ALL" 2>&1 | cut -c1-220
