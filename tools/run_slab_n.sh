N=$1; SIZE=$2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N tools/slab_probe.py $SIZE 8 > gpurun_out/slab_${N}_${SIZE}.log 2>&1; grep -B2 -A8 "Traceback" gpurun_out/slab_${N}_${SIZE}.log | head -20; grep "slab vs\|rep " gpurun_out/slab_${N}_${SIZE}.log
