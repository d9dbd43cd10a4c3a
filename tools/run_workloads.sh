python bench.py --workload c2 --steps 2 --warmup 1 > gpurun_out/bench_c2_n1.json 2> gpurun_out/bench_c2_n1.err; tail -2 gpurun_out/bench_c2_n1.err; cut -c1-400 gpurun_out/bench_c2_n1.json
python bench.py --workload c4 --items 512 --steps 2 --warmup 1 > gpurun_out/bench_c4_n1.json 2> gpurun_out/bench_c4_n1.err; tail -2 gpurun_out/bench_c4_n1.err; cut -c1-400 gpurun_out/bench_c4_n1.json
python bench.py --workload c5 --size 256 --steps 2 --warmup 1 > gpurun_out/bench_c5_n1.json 2> gpurun_out/bench_c5_n1.err; tail -2 gpurun_out/bench_c5_n1.err; cut -c1-400 gpurun_out/bench_c5_n1.json
