#!/bin/bash
# A/B of the TMA stencil's stage count / CTAs per SM (rebuilds stencil.o on the box)
cd lsml_b200/csrc
for v in "-DTM_STAGES_N=4 -DTM_MIN_CTAS=3" "-DTM_STAGES_N=6 -DTM_MIN_CTAS=3" "-DTM_STAGES_N=4 -DTM_MIN_CTAS=4" "-DTM_STAGES_N=3 -DTM_MIN_CTAS=4" "-DTM_STAGES_N=8 -DTM_MIN_CTAS=2"; do
    rm -f build/stencil.o
    make EXTRA="$v -Xptxas -v" 2>&1 | grep -A2 "gmag_os_tma_kernelILb1ELb1" | grep -E "registers|spill" | head -2
    echo "== $v"
    (cd ../..; timeout 200 python tools/stencil_probe.py 256 2>&1 | grep -E "TMA, NULL|gmag_os f64 dx=1  " ; timeout 100 python tools/stencil_probe.py 512 2>&1 | grep -E "gmag_os f64 dx=1  ")
done
