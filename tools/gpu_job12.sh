#!/bin/bash
mkdir -p gpurun_out
python tools/check_env_bitwise.py C4 NE_B200_TAB2_SHAPE=0 NE_B200_TAB2_SHAPE=1 > gpurun_out/j12_bitwise.log 2>&1
python tools/check_env_bitwise.py C2 NE_B200_TAB2_SHAPE=0 NE_B200_TAB2_SHAPE=2 >> gpurun_out/j12_bitwise.log 2>&1
timeout 900 python tools/time_ao.py --config C4 --out j12_time_ao "" NE_B200_TAB2_SHAPE=1 NE_B200_TAB2_SHAPE=2 "" NE_B200_TAB2_SHAPE=1 NE_B200_TAB2_SHAPE=2 > gpurun_out/j12_time_ao.log 2>&1
NE_B200_TAB2_SHAPE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:ao_flux_tab2 -s 2 -c 1 -f -o gpurun_out/j12_tab2_shape1 python tools/prof_ao.py C4 f64 > gpurun_out/j12_ncu.log 2>&1
cat gpurun_out/j12_bitwise.log gpurun_out/j12_time_ao.log
