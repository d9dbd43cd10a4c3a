set -e
python tools/iter_probe.py 256 8 > gpurun_out/iter_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python tools/iter_probe.py 256 8 > gpurun_out/ncu_iter.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:fmm_march_kernel --launch-skip 12 --launch-count 1 -o gpurun_out/fmm_march_r1 -f python tools/iter_probe.py 256 8 > gpurun_out/ncu_fmm.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:forest_predict --launch-skip 10 --launch-count 1 -o gpurun_out/forest_r1 -f python tools/iter_probe.py 256 8 > gpurun_out/ncu_forest.log 2>&1
python tools/stencil_probe.py 256 > gpurun_out/stencil_probe.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:gmag_os_march --launch-skip 3 --launch-count 1 -o gpurun_out/stencil_r1 -f python tools/stencil_probe.py 256 3 > gpurun_out/ncu_stencil.log 2>&1
cat gpurun_out/iter_plain.log gpurun_out/stencil_probe.log
