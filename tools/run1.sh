python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/fmm_probe.py 256 12 > gpurun_out/fmm_probe.log 2>&1; tail -13 gpurun_out/fmm_probe.log
