#!/bin/bash
# Round 2, GPU call 22 (one GPU): how small must a work set be for the single-CTA tail sweeps to beat a
# grid-wide sweep?  (tail entries dealt one per warp first)
cd lsml_b200/csrc
for v in "" "-DFMM_TAIL_MAX_ENTRIES=32" "-DFMM_TAIL_MAX_ENTRIES=8" "-DFMM_TAIL_MAX_ENTRIES=0"; do
    rm -f build/fmm.o
    make EXTRA="$v" 2>&1 | grep -E "error" | head -3
    echo "== ${v:-default (256)}"
    (cd ../..; timeout 200 python tools/fmm_probe.py 256 12 2>&1 | grep -E "^it  (5|6|7|8|9|11)" | cut -c50-200)
done
rm -f build/fmm.o; make 2>&1 | grep -E "error" | head -3
