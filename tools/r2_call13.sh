#!/bin/bash
# Round 2, GPU call 13 (one GPU): the tensor-core MLP kernel and the cached host-path buffers.
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_engine_gpu.py tests/test_stencil_gpu.py tests/test_model_gpu.py -m gpu -x -q 2>&1 | tail -n 8
timeout 200 python tools/mlp_probe.py 2>&1 | tail -n 1
LSML_MLP_VARIANT=scalar timeout 200 python tools/mlp_probe.py 2>&1 | tail -n 1
