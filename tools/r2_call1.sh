#!/bin/bash
# Round 2, GPU call 1 (one GPU): the suite, smoke, the C3 bench, and the C4 bisect under short
# per-step timeouts (bench.py's own watchdog dumps the host stack and the live march state).
mkdir -p gpurun_out
O=gpurun_out
run() { # name, seconds, command...
    local name=$1 secs=$2; shift 2
    echo "=== $name"
    timeout $secs "$@" > $O/$name.out 2> $O/$name.err
    echo "exit $? ($name)"; tail -n 2 $O/$name.out | cut -c1-600
}
run tests_gpu 900 python -m pytest tests -m gpu -x -q
run smoke 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
for items in 512 1024 2048; do
    run bench_c4_$items 300 python bench.py --workload c4 --items $items --steps 1 --warmup 1 --verbose --watchdog 240
    tail -n 4 $O/bench_c4_$items.err | cut -c1-300
done
run bench_c3 400 python bench.py --steps 2 --warmup 3 --watchdog 360
run bench_c2 240 python bench.py --workload c2 --steps 2 --warmup 1 --watchdog 200
