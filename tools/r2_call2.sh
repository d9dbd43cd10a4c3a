#!/bin/bash
# Round 2, GPU call 2 (one GPU): suite (incl. the BASELINE-size parity tests), smoke, the default
# bench with its C4 / C5 sections at N = 1, C2.
mkdir -p gpurun_out
O=gpurun_out
run() { # name, seconds, command...
    local name=$1 secs=$2; shift 2
    echo "=== $name"
    timeout $secs "$@" > $O/$name.out 2> $O/$name.err
    echo "exit $? ($name)"; tail -n 2 $O/$name.out | cut -c1-1500
}
run tests_gpu 1500 python -m pytest tests -m gpu -x -q --durations=8
tail -n 30 $O/tests_gpu.out | cut -c1-300
run smoke 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
run bench_default 900 python bench.py --steps 3 --warmup 3 --verbose --watchdog 800
tail -n 25 $O/bench_default.err | cut -c1-300
run bench_c2 240 python bench.py --workload c2 --watchdog 200
run bench_ref 400 python bench.py --impl reference --steps 2 --warmup 1
