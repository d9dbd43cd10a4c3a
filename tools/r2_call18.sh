#!/bin/bash
# Round 2, GPU call 18 (one GPU): register-march Gaussian for the strided axes (planes bit-exact vs scipy
# in the suite), timing against the tap-gather kernel.
mkdir -p gpurun_out
O=gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > $O/r2c18_tests.out 2> $O/r2c18_tests.err
echo "exit $?"; tail -n 3 $O/r2c18_tests.out | cut -c1-300
timeout 200 python tools/planes_probe.py 256 2>&1 | tail -n 1
LSML_GAUSS_NO_MARCH=1 timeout 200 python tools/planes_probe.py 256 2>&1 | tail -n 1
timeout 200 python bench.py --workload c4 --items 2048 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4-2048', d['ms_per_step'], d['phases_ms_per_engine_batch'])"
