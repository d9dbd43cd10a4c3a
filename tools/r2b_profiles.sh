#!/bin/bash
# Round 2 (second session) profiling pass (one GPU), same protocol as tools/r2_profiles.sh: each
# program first runs plainly (exit 0 without ncu), then under ncu.  Outputs land in gpurun_out/;
# tools/r2b_profiles_summarise.sh turns them into the committed summaries under profiles/.
set -x
mkdir -p gpurun_out
O=gpurun_out
python tools/iter_probe.py 256 8 > $O/r2b_iter_plain.log 2>&1 || exit 1
python tools/forest_probe.py > $O/r2b_forest_probe.log 2>&1 || exit 1
python tools/fmm_probe.py 256 12 > $O/r2b_fmm_probe.log 2>&1 || exit 1
# launch list of the same command (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2b_launches_iter_probe.csv \
    python tools/iter_probe.py 256 8 > $O/r2b_ncu_iter.log 2>&1
cap() { # name, kernel regex, skip, program...
    local name=$1 rx=$2 skip=$3; shift 3
    ncu --set full --import-source on --clock-control none -k regex:$rx --launch-skip $skip --launch-count 1 \
        -o $O/$name -f "$@" > $O/ncu_$name.log 2>&1
}
cap r2b_fmm_march fmm_march_kernel 12 python tools/iter_probe.py 256 8
cap r2b_forest_iter forest_ranked_kernel 10 python tools/iter_probe.py 256 8
cap r2b_forest_ranked forest_ranked_kernel 5 python tools/forest_probe.py
cap r2b_band_update band_update_kernel 10 python tools/iter_probe.py 256 8
cap r2b_band_apply band_apply_kernel 10 python tools/iter_probe.py 256 8
cat $O/r2b_iter_plain.log $O/r2b_forest_probe.log; tail -n 4 $O/r2b_fmm_probe.log | cut -c1-300
