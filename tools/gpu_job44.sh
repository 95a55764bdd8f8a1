#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"interp_rows_kernel|post_solve_kernel|trip_order_histogram" -s 6 -c 3 -f -o gpurun_out/j44_step python tools/prof_step.py > gpurun_out/j44_ncu.log 2>&1
ls -la gpurun_out/j44_step.ncu-rep; tail -3 gpurun_out/j44_ncu.log
