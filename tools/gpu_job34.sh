#!/bin/bash
mkdir -p gpurun_out
for n in 8 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 40 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/j34_bench_n$n.json 2> gpurun_out/j34_bench_n$n.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2962$n bench.py --gpus $n --steps 40 --warmup 5 --no-cpu-baseline --no-extras --no-time-rebalance > gpurun_out/j34_bench_n${n}_notime.json 2> gpurun_out/j34_bench_n${n}_notime.err
done
tail -3 gpurun_out/j34_bench_n8.err | cut -c1-300
