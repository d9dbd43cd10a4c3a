#!/bin/bash
# Round-2 profiling pass (one GPU).  Each program first runs plainly (exit 0 without ncu), then
# under ncu.  Outputs land in gpurun_out/; tools/r2_profiles_summarise.sh turns them into the
# committed summaries under profiles/.
set -x
mkdir -p gpurun_out
O=gpurun_out
python tools/iter_probe.py 256 8 > $O/r2_iter_plain.log 2>&1 || exit 1
python tools/stencil_probe.py 256 > $O/r2_stencil_probe.log 2>&1 || exit 1
# launch list of the same command (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_iter_probe.csv \
    python tools/iter_probe.py 256 8 > $O/r2_ncu_iter.log 2>&1
# one full capture per kernel of interest
cap() { # name, kernel regex, skip, program...
    local name=$1 rx=$2 skip=$3; shift 3
    ncu --set full --import-source on --clock-control none -k regex:$rx --launch-skip $skip --launch-count 1 \
        -o $O/$name -f "$@" > $O/ncu_$name.log 2>&1
}
cap r2_fmm_march fmm_march_kernel 12 python tools/iter_probe.py 256 8
cap r2_forest forest_predict_kernel 10 python tools/iter_probe.py 256 8
cap r2_band_update band_update_kernel 10 python tools/iter_probe.py 256 8
cap r2_band_apply band_apply_kernel 10 python tools/iter_probe.py 256 8
cap r2_stencil_tma gmag_os_tma_kernel 20 python tools/stencil_probe.py 256
cat $O/r2_iter_plain.log $O/r2_stencil_probe.log
