ncu --set full --import-source on --clock-control none -k regex:gmag_os_march --launch-skip 3 --launch-count 1 -o gpurun_out/stencil_r1v -f python tools/stencil_probe.py 256 3 > gpurun_out/ncu_stencil.log 2>&1
tail -2 gpurun_out/ncu_stencil.log
