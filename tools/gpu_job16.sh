#!/bin/bash
mkdir -p gpurun_out
python tools/time_ao.py --config C1 --out j16_c1 "" NE_B200_TAB2_NO_ORDER=1 > gpurun_out/j16.log 2>&1
python tools/time_ao.py --config C2 --out j16_c2 "" NE_B200_TAB2_NO_ORDER=1 >> gpurun_out/j16.log 2>&1
cat gpurun_out/j16.log
