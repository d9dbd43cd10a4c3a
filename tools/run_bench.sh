python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -2 gpurun_out/bench_r1.err; cat gpurun_out/bench_r1.json
