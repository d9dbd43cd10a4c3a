#!/bin/bash
# bench.py exactly as the driver launches it for N GPUs (SCALE_rNN): tools/r2_scale.sh N
N=${1:-2}
mkdir -p gpurun_out
timeout 1400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --verbose --watchdog 1200 \
    > gpurun_out/scale_n$N.out 2> gpurun_out/scale_n$N.err
echo "exit $?"; tail -n 1 gpurun_out/scale_n$N.out | cut -c1-300; grep -E "section|c4:|c5:" gpurun_out/scale_n$N.err | tail -12
