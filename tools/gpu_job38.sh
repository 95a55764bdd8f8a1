#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flux_queue_kernel -s 2 -c 1 -f -o gpurun_out/j38_asi python tools/prof_seaice.py C3 > gpurun_out/j38_ncu.log 2>&1
ls -la gpurun_out/j38_asi.ncu-rep
