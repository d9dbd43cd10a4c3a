#!/bin/bash
# First GPU call of round 2 (one GPU, ~10 min):  gpurun --timeout 1500 -- 'bash tools/run_round2_first.sh'
# Everything changed on the CPU after round 1's GPU budget ran out is exercised here, each step
# under its own short timeout so that a hang costs minutes, not the round:
#   contributor rule (oracle + kernels), hint tags 1..255 / wipe, total sweep limit, COMRaySample,
#   SlabLayout refactor, the C4 batch sizes up to the 2048 crops that did not finish in round 1.
mkdir -p gpurun_out
O=gpurun_out
run() { # name, seconds, command...
    local name=$1 secs=$2; shift 2
    echo "=== $name"
    timeout $secs "$@" > $O/$name.out 2> $O/$name.err
    echo "exit $? ($name)"; tail -n 3 $O/$name.out | cut -c1-400
}
run tests_gpu 900 python -m pytest tests -m gpu -x -q
run smoke 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
run bench_c3 600 python bench.py --steps 3 --warmup 3
for items in 256 512 1024 2048; do
    run bench_c4_$items 240 python bench.py --workload c4 --items $items --steps 1 --warmup 1
done
run bench_c2 240 python bench.py --workload c2 --steps 2 --warmup 1
run bench_c2_f32 240 python bench.py --workload c2 --features gestalt32 --steps 2 --warmup 1
run slab_512 300 python tools/slab_probe.py 512
