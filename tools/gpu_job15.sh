#!/bin/bash
mkdir -p gpurun_out
python tools/time_band.py 8 3 NE_B200_TAB2_LPT=0 "" NE_B200_TAB2_LPT=0 "" > gpurun_out/j15_band.log 2>&1
python tools/time_band.py 8 0 NE_B200_TAB2_LPT=0 "" >> gpurun_out/j15_band.log 2>&1
python tools/time_band.py 4 1 NE_B200_TAB2_LPT=0 "" >> gpurun_out/j15_band.log 2>&1
python tools/time_band.py 1 0 NE_B200_TAB2_LPT=100000 "" >> gpurun_out/j15_band.log 2>&1
python tools/check_env_bitwise.py C2 NE_B200_TAB2_LPT=0 NE_B200_TAB2_LPT=100000 >> gpurun_out/j15_band.log 2>&1
cat gpurun_out/j15_band.log
