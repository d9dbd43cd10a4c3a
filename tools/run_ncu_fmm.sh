ncu --set full --import-source on --clock-control none -k regex:fmm_march_kernel --launch-skip 10 --launch-count 1 -o gpurun_out/fmm_march_r1 -f python tools/fmm_probe.py 256 6 > gpurun_out/ncu_fmm.log 2>&1
tail -3 gpurun_out/ncu_fmm.log
