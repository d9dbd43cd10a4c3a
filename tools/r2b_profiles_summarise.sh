#!/bin/bash
# gpurun_out/ (scratch) -> profiles/ (committed), second session of round 2: launch-list summary,
# per-kernel ncu summaries, the traffic file bench.py reads (entries of kernels that did not change
# since tools/r2_profiles.sh -- the TMA stencil -- are kept).
O=gpurun_out
P=profiles
cp $O/r2b_launches_iter_probe.csv $P/r2b_launches_iter_probe.csv
python tools/summarize_launches.py $O/r2b_launches_iter_probe.csv > $P/r2b_launches_iter_probe.txt
python tools/ncu_summary.py $O/r2b_fmm_march.ncu-rep --json fmm_march_kernel $P/r2_ncu_traffic.json > $P/r2b_ncu_fmm_march.txt
python tools/ncu_summary.py $O/r2b_forest_iter.ncu-rep --json forest_ranked_kernel $P/r2_ncu_traffic.json > $P/r2b_ncu_forest_ranked_iter.txt
python tools/ncu_summary.py $O/r2b_forest_ranked.ncu-rep > $P/r2b_ncu_forest_ranked.txt
python tools/ncu_summary.py $O/r2b_band_update.ncu-rep --json band_update_kernel $P/r2_ncu_traffic.json > $P/r2b_ncu_band_update.txt
python tools/ncu_summary.py $O/r2b_band_apply.ncu-rep --json band_apply_kernel $P/r2_ncu_traffic.json > $P/r2b_ncu_band_apply.txt
cp $O/r2b_iter_plain.log $P/r2b_iter_probe.txt
cp $O/r2b_forest_probe.log $P/r2b_forest_probe.txt
cp $O/r2b_fmm_probe.log $P/r2b_fmm_probe.txt
cat $P/r2_ncu_traffic.json
