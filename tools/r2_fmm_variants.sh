#!/bin/bash
# A/B of march-kernel build variants (rebuilds fmm.o on the box): tools/r2_fmm_variants.sh "-DX=.." ...
cd lsml_b200/csrc
for v in "$@"; do
    rm -f build/fmm.o
    make EXTRA="$v" 2>&1 | grep -E "error" | head -3
    echo "== $v"
    (cd ../..; timeout 200 python tools/fmm_probe.py 256 12 2>&1 | grep -E "^it  (5|8|11)|^phase" | cut -c1-200; timeout 200 python bench.py --workload c4 --items 2048 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4-2048', d['ms_per_step'], d['phases_ms_per_engine_batch']['rebuild'])")
done
rm -f build/fmm.o; make 2>&1 | grep -E "error" | head -3
