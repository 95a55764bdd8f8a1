#!/bin/bash
mkdir -p gpurun_out
python tools/check_env_bitwise.py C4 "" "NE_B200_TAB2_ORDER=0" > gpurun_out/j28_bitwise.log 2>&1
python tools/check_env_bitwise.py C2 "" "NE_B200_TAB2_NO_ORDER=1" >> gpurun_out/j28_bitwise.log 2>&1
grep bitwise gpurun_out/j28_bitwise.log
python tools/time_ao.py --out j28_time_ao "" NE_B200_TAB2_ORDER=1 NE_B200_TAB2_ORDER=0 "" > gpurun_out/j28_time_ao.log 2>&1
cat gpurun_out/j28_time_ao.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"trip_order" -c 3 --csv --log-file gpurun_out/j28_hist.csv python tools/prof_ao.py C4 f64 > gpurun_out/j28_a.log 2>&1
grep -E "gpu__time_duration" gpurun_out/j28_hist.csv | tail -2 | sed 's/.*"\(void [a-z_0-9]*\).*Command line profiler metrics",/\1 /'
NE_CFG=C4 python tools/time_step.py "" > gpurun_out/j28_step_C4.log 2>&1; tail -1 gpurun_out/j28_step_C4.log
