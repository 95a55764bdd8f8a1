#!/bin/bash
mkdir -p gpurun_out
for c in C1 C2; do NE_CFG=$c python tools/time_step.py "" > gpurun_out/j32_step_$c.log 2>&1; tail -1 gpurun_out/j32_step_$c.log; done
timeout 600 ncu --nvtx --print-nvtx-rename kernel --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/j32_seaice_nvtx.csv python tools/time_seaice.py > gpurun_out/j32_seaice.log 2>&1
tail -3 gpurun_out/j32_seaice.log
