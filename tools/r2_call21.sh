#!/bin/bash
# Round 2, GPU call 21 (one GPU): dense stencils with arbitrary spacing -- the quotient by FMA instead of the
# division sequence (bit-exact vs the compiled reference in the suite), timings.
mkdir -p gpurun_out
O=gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > $O/r2c21_tests.out 2> $O/r2c21_tests.err
echo "exit $?"; tail -n 3 $O/r2c21_tests.out | cut -c1-300
timeout 300 python tools/stencil_probe.py 256 2>&1 | tail -n 16 | cut -c1-200
