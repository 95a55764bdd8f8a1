#!/bin/bash
mkdir -p gpurun_out
for ord in 2 1; do
NE_B200_TAB2_ORDER=$ord timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python tools/prof_ao.py C1 f64 > gpurun_out/j36_racecheck_order$ord.log 2>&1; tail -2 gpurun_out/j36_racecheck_order$ord.log
NE_B200_TAB2_ORDER=$ord timeout 600 compute-sanitizer --tool memcheck python tools/prof_ao.py C1 f64 > gpurun_out/j36_memcheck_order$ord.log 2>&1; tail -1 gpurun_out/j36_memcheck_order$ord.log
done
timeout 600 compute-sanitizer --tool memcheck python tools/time_seaice.py C1 > gpurun_out/j36_memcheck_seaice.log 2>&1; tail -1 gpurun_out/j36_memcheck_seaice.log
