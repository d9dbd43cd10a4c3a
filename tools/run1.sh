set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/stencil_probe.py 256 > gpurun_out/stencil_probe.log 2>&1; cat gpurun_out/stencil_probe.log
python tools/fmm_probe.py 256 12 > gpurun_out/fmm_probe.log 2>&1; tail -15 gpurun_out/fmm_probe.log
