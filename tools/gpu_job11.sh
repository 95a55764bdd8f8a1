#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/j11_bench.json 2> gpurun_out/j11_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/j11_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --sustained-seconds 0 --no-parity > gpurun_out/j11_launches_bench.log 2>&1
timeout 300 python tools/time_seaice.py > gpurun_out/j11_time_seaice.log 2>&1
tail -3 gpurun_out/j11_bench.err
