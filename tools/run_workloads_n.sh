N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N"
$TR bench.py --gpus $N --workload c5 --size 1024 --steps 1 --warmup 1 > gpurun_out/bench_c5_1024_n$N.json 2> gpurun_out/bench_c5_1024_n$N.err; grep -A6 Traceback gpurun_out/bench_c5_1024_n$N.err | head; cut -c1-330 gpurun_out/bench_c5_1024_n$N.json
$TR bench.py --gpus $N --workload c4 --items 2048 --steps 1 --warmup 1 > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err; grep -A6 Traceback gpurun_out/bench_c4_n$N.err | head; cut -c1-330 gpurun_out/bench_c4_n$N.json
