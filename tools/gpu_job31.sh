#!/bin/bash
mkdir -p gpurun_out
for b in "8 3" "8 0" "4 1" "1 0"; do python tools/time_band.py $b "" NE_B200_TAB2_DESCENDING=1 "" NE_B200_TAB2_DESCENDING=1; done > gpurun_out/j31_band.log 2>&1
cat gpurun_out/j31_band.log
python tools/check_env_bitwise.py C2 "" "NE_B200_TAB2_DESCENDING=1" 2>&1 | grep bitwise
