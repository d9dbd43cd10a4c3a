#!/bin/bash
# gpurun_out/ (scratch) -> profiles/ (committed): launch-list summary, per-kernel ncu summaries and
# the traffic file bench.py reads.
O=gpurun_out
P=profiles
cp $O/r2_launches_iter_probe.csv $P/r2_launches_iter_probe.csv
python tools/summarize_launches.py $O/r2_launches_iter_probe.csv > $P/r2_launches_iter_probe.txt
rm -f $P/r2_ncu_traffic.json
python tools/ncu_summary.py $O/r2_fmm_march.ncu-rep --json fmm_march_kernel $P/r2_ncu_traffic.json > $P/r2_ncu_fmm_march.txt
python tools/ncu_summary.py $O/r2_forest.ncu-rep --json forest_predict_kernel $P/r2_ncu_traffic.json > $P/r2_ncu_forest_iter.txt
python tools/ncu_summary.py $O/r2_band_update.ncu-rep --json band_update_kernel $P/r2_ncu_traffic.json > $P/r2_ncu_band_update.txt
python tools/ncu_summary.py $O/r2_band_apply.ncu-rep --json band_apply_kernel $P/r2_ncu_traffic.json > $P/r2_ncu_band_apply.txt
python tools/ncu_summary.py $O/r2_stencil_tma.ncu-rep --json gmag_os_tma_kernel $P/r2_ncu_traffic.json > $P/r2_ncu_stencil_tma.txt
cp $O/r2_iter_plain.log $P/r2_iter_probe.txt
cp $O/r2_stencil_probe.log $P/r2_stencil_probe.txt
cat $P/r2_ncu_traffic.json
