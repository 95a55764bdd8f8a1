#!/bin/bash
mkdir -p gpurun_out
python tools/time_interp.py C4 "" NE_B200_INTERP_ROWS=2 NE_B200_INTERP_ROWS=4 NE_B200_INTERP_ROWS=8 "" NE_B200_INTERP_ROWS=4 > gpurun_out/j41_interp.log 2>&1
cat gpurun_out/j41_interp.log
