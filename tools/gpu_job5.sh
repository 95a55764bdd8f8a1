#!/bin/bash
mkdir -p gpurun_out
python tools/check_env_bitwise.py C4 NE_B200_TAB2_PPT=1 NE_B200_TAB2_PPT=2 > gpurun_out/j5_bitwise.log 2>&1
python tools/check_env_bitwise.py C2 NE_B200_TAB2_PPT=1,NE_B200_TAB2_NO_ORDER=1 NE_B200_TAB2_PPT=2 >> gpurun_out/j5_bitwise.log 2>&1
timeout 900 python tools/time_ao.py --config C4 --out j5_time_ao NE_B200_TAB2_PPT=1 "" NE_B200_TAB2_WINDOW=512 NE_B200_TAB2_NO_ORDER=1 NE_B200_TAB_WAVES=4 NE_B200_TAB_WAVES=8 NE_B200_TAB2_PPT=1 "" > gpurun_out/j5_time_ao.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ao_flux_tab2 -s 2 -c 1 -f -o gpurun_out/j5_tab2n python tools/prof_ao.py C4 f64 > gpurun_out/j5_ncu.log 2>&1
cat gpurun_out/j5_bitwise.log gpurun_out/j5_time_ao.log
