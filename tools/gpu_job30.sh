#!/bin/bash
mkdir -p gpurun_out
for n in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/j30_bench_n$n.json 2> gpurun_out/j30_bench_n$n.err
done
timeout 600 python bench.py --gpus 1 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/j30_bench_n1.json 2> gpurun_out/j30_bench_n1.err
tail -2 gpurun_out/j30_bench_n8.err
