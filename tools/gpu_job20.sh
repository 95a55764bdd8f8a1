#!/bin/bash
mkdir -p gpurun_out
for c in C1 C2 C4; do NE_CFG=$c python tools/time_step.py "" > gpurun_out/j20_step_$c.log 2>&1; done
python tools/time_seaice.py > gpurun_out/j20_seaice.log 2>&1
tail -n 3 gpurun_out/j20_step_C1.log gpurun_out/j20_step_C2.log gpurun_out/j20_step_C4.log; tail -4 gpurun_out/j20_seaice.log
