#!/bin/bash
# A/B of the march kernel's occupancy target (rebuilds fmm.o on the box)
cd lsml_b200/csrc
for v in "-DFMM_MIN_BLOCKS=4" "-DFMM_MIN_BLOCKS=3" "-DFMM_MIN_BLOCKS=5" "-DFMM_MIN_BLOCKS=6" "-DFMM_MIN_BLOCKS=2"; do
    rm -f build/fmm.o
    make EXTRA="$v -Xptxas -v" 2>&1 | grep -A2 "fmm_march_kernelILi3" | grep -E "registers|spill" | head -2
    echo "== $v"
    (cd ../..; timeout 200 python tools/fmm_probe.py 256 12 2>&1 | grep -E "^it  (5|8|11)")
done
