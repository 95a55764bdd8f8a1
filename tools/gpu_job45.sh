#!/bin/bash
mkdir -p gpurun_out
python tools/time_interp.py C4 NE_B200_INTERP_ROWS=1 "" NE_B200_INTERP_ROWS=8 "" > gpurun_out/j45_interp.log 2>&1
cat gpurun_out/j45_interp.log
