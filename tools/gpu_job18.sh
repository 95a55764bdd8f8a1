#!/bin/bash
mkdir -p gpurun_out
python tools/time_interp.py C4 NE_B200_INTERP_STAGED_V1=1 NE_B200_INTERP_TILE_ROWS=1 NE_B200_INTERP_TILE_ROWS=2 NE_B200_INTERP_TILE_ROWS=4 > gpurun_out/j18_interp.log 2>&1
cat gpurun_out/j18_interp.log
