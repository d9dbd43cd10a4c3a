python tools/iter_probe.py 256 8 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_iter.csv python tools/iter_probe.py 256 8 > gpurun_out/ncu_iter.log 2>&1
tail -2 gpurun_out/ncu_iter.log
